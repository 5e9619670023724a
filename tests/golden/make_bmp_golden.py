"""BMP textures, decoded by the REFERENCE'S OWN loader (scene.cpp -> stb_image, oracle/_ref/ref_cpu --b2s).

PIL writes 1-bit, 8-bit grey, 8-bit palette, 24-bit and 32-bit files; hand-assembled ones add a 4-bit palette,
16-bit RGB555 (BI_RGB) and RGB565 (BI_BITFIELDS), a top-down 24-bit file (negative height), an OS/2 header
(12 bytes) and a 32-bit file whose fourth byte is zero everywhere (stb makes it opaque).  Each becomes the map_Kd of
tests/golden/quadbox.obj and goes through the reference's loader.  bmp/<name>.bmp is the input,
bmp/texels.npz[<name>] the texels the reference holds after loading (H x W x C, rows flipped as scene.cpp:133 does).
Every file holds at least 6168 bytes of texels (scene.cpp:148 prints data[2055*3 ...]).
Needs /root/reference (through oracle/_ref) and PIL; the outputs are committed.
"""
import os
import shutil
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import harness  # noqa: E402
from mygpuraytracer_b200 import scenes  # noqa: E402
from mygpuraytracer_b200.podscene import PodScene  # noqa: E402


def file_header(offset, size):
    return b"BM" + struct.pack("<IHHI", size, 0, 0, offset)


def info40(w, h, bpp, compress=0, colours=0):
    return struct.pack("<IiiHHIIiiII", 40, w, h, 1, bpp, compress, 0, 2835, 2835, colours, 0)


def rows(data, pad_to=4):
    out = b""
    for r in data:
        r = bytes(r)
        out += r + b"\0" * ((-len(r)) % pad_to)
    return out


def write_files(dst):
    from PIL import Image

    rng = np.random.default_rng(0xB3)

    def rgb(w, h, noise=6):
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        img = np.stack([127 + 110 * np.sin(x * 0.23 + y * 0.15), 127 + 110 * np.cos(y * 0.21 - x * 0.12),
                        127 + 110 * np.sin((x + y) * 0.18)], -1) + rng.normal(0, noise, (h, w, 3))
        return np.clip(img, 0, 255).astype(np.uint8)

    names = []
    pil = {
        "rgb24": Image.fromarray(rgb(91, 73)),
        "rgba32": Image.fromarray(np.dstack([rgb(85, 77), rng.integers(1, 256, (77, 85), dtype=np.uint8)]), "RGBA"),
        "grey8": Image.fromarray(rgb(93, 79)[..., 0], "L"),
        "palette8": Image.fromarray(rgb(89, 81)).quantize(colors=150),
        "onebit": Image.fromarray((rgb(101, 83)[..., 0] > 127).astype(np.uint8) * 255).convert("1"),
    }
    for name, img in pil.items():
        img.save(os.path.join(dst, name + ".bmp"), "BMP")
        names.append(name)

    def put(name, blob):
        with open(os.path.join(dst, name + ".bmp"), "wb") as f:
            f.write(blob)
        names.append(name)

    # 4-bit palette, bottom-up
    w, h = 87, 75
    idx = rng.integers(0, 16, (h, w), dtype=np.uint8)
    pal = rng.integers(0, 256, (16, 4), dtype=np.uint8)
    pal[:, 3] = 0
    packed = [np.packbits(np.unpackbits(r[:, None], axis=1)[:, 4:].reshape(-1)) for r in idx]
    data = rows(packed)
    off = 14 + 40 + 64
    put("palette4", file_header(off, off + len(data)) + info40(w, h, 4, colours=16) + pal.tobytes() + data)
    # 16-bit RGB555 (BI_RGB) and RGB565 (BI_BITFIELDS)
    w, h = 82, 76
    px = rng.integers(0, 1 << 16, (h, w)).astype("<u2")
    data = rows([r.tobytes() for r in px])
    off = 14 + 40
    put("rgb555", file_header(off, off + len(data)) + info40(w, h, 16) + data)
    off = 14 + 40 + 12
    put("rgb565", file_header(off, off + len(data)) + info40(w, h, 16, compress=3) +
        struct.pack("<III", 0xF800, 0x07E0, 0x001F) + data)
    # 24-bit, top-down (negative height)
    w, h = 83, 77
    img = rgb(w, h)
    data = rows([r[:, ::-1].tobytes() for r in img])
    off = 14 + 40
    put("rgb24_topdown", file_header(off, off + len(data)) + info40(w, -h, 24) + data)
    # OS/2 header (12 bytes), 24-bit, bottom-up
    w, h = 81, 79
    img = rgb(w, h)
    data = rows([r[:, ::-1].tobytes() for r in img[::-1]])
    off = 14 + 12
    put("os2_rgb24", file_header(off, off + len(data)) + struct.pack("<IHHHH", 12, w, h, 1, 24) + data)
    # 32-bit BI_RGB whose fourth byte is 0 everywhere
    w, h = 80, 78
    img = np.dstack([rgb(w, h)[..., ::-1], np.zeros((h, w), np.uint8)])
    data = rows([r.tobytes() for r in img[::-1]])
    off = 14 + 40
    put("rgbx32_zero_alpha", file_header(off, off + len(data)) + info40(w, h, 32) + data)
    return names


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "bmp")
    os.makedirs(dst, exist_ok=True)
    tex_dir = os.path.join(os.path.dirname(harness.RUN_MODELS), "textures")
    os.makedirs(tex_dir, exist_ok=True)
    obj = open(os.path.join(HERE, "quadbox.obj")).read()
    texels = {}
    for name in write_files(dst):
        shutil.copyfile(os.path.join(dst, name + ".bmp"), os.path.join(tex_dir, f"bm_{name}.bmp"))
        with open(os.path.join(harness.RUN_MODELS, f"bm_{name}.obj"), "w") as f:
            f.write(obj.replace("mtllib quadbox.mtl", f"mtllib bm_{name}.mtl"))
        with open(os.path.join(harness.RUN_MODELS, "materials", f"bm_{name}.mtl"), "w") as f:
            f.write(f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/bm_{name}.bmp\n")
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        with open(txt, "w") as f:
            f.write(scenes.scene_text("cornellObj", width=16, height=16, obj_path=f"../models/bm_{name}.obj"))
        b2s = os.path.join(d, "s.b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
        ref = PodScene.load(b2s)
        assert len(ref.textures) == 1, (name, len(ref.textures))
        texels[name] = ref.textures[0]
        print(name, ref.textures[0].shape, os.path.getsize(os.path.join(dst, name + ".bmp")))
        shutil.rmtree(d)
    np.savez_compressed(os.path.join(dst, "texels.npz"), **texels)


if __name__ == "__main__":
    main()

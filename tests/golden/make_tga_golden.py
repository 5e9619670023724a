"""TGA textures, decoded by the REFERENCE'S OWN loader (scene.cpp -> stb_image, oracle/_ref/ref_cpu --b2s).

PIL writes grey, RGB, RGBA and colour-mapped files, raw and run-length encoded, stored bottom-up and top-down;
a 16-bit RGB555 file is assembled by hand (PIL cannot write one).  Each becomes the map_Kd of
tests/golden/quadbox.obj and goes through the reference's loader.  tga/<name>.tga is the input,
tga/texels.npz[<name>] the texels the reference holds after loading (H x W x C, rows flipped as scene.cpp:133 does).
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


def write_files(dst):
    from PIL import Image

    rng = np.random.default_rng(0x76A)

    def rgb(w, h, noise=6):
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        img = np.stack([127 + 110 * np.sin(x * 0.27 + y * 0.11), 127 + 110 * np.cos(y * 0.19 - x * 0.13),
                        127 + 110 * np.sin((x + y) * 0.21)], -1) + rng.normal(0, noise, (h, w, 3))
        return np.clip(img, 0, 255).astype(np.uint8)

    flat = rgb(90, 76, noise=0)
    flat[20:50, 10:60] = (200, 40, 90)  # long runs for the RLE files
    cases = {
        "rgb_raw_bottomup": (Image.fromarray(rgb(91, 73)), dict(orientation=-1)),
        "rgb_raw_topdown": (Image.fromarray(rgb(91, 73)), dict(orientation=1)),
        "rgb_rle": (Image.fromarray(flat), dict(compression="tga_rle")),
        "rgba_raw": (Image.fromarray(np.dstack([rgb(84, 78), rng.integers(0, 256, (78, 84), dtype=np.uint8)]), "RGBA"), {}),
        "rgba_rle_topdown": (Image.fromarray(np.dstack([flat, np.full(flat.shape[:2], 130, np.uint8)]), "RGBA"),
                             dict(compression="tga_rle", orientation=1)),
        "grey_raw": (Image.fromarray(rgb(95, 81)[..., 0], "L"), {}),
        "grey_rle": (Image.fromarray(flat[..., 1].copy(), "L"), dict(compression="tga_rle")),
        "palette_raw": (Image.fromarray(rgb(88, 80)).quantize(colors=120), {}),
        "palette_rle": (Image.fromarray(flat).quantize(colors=40), dict(compression="tga_rle")),
    }
    names = []
    for name, (img, opts) in cases.items():
        img.save(os.path.join(dst, name + ".tga"), "TGA", **opts)
        names.append(name)
    # 16-bit true colour (RGB555), bottom-up, raw
    w, h = 80, 78
    px = rng.integers(0, 1 << 15, (h, w)).astype("<u2")
    with open(os.path.join(dst, "rgb555_raw.tga"), "wb") as f:
        f.write(struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, w, h, 16, 0))
        f.write(px.tobytes())
    names.append("rgb555_raw")
    return names


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "tga")
    os.makedirs(dst, exist_ok=True)
    tex_dir = os.path.join(os.path.dirname(harness.RUN_MODELS), "textures")
    os.makedirs(tex_dir, exist_ok=True)
    obj = open(os.path.join(HERE, "quadbox.obj")).read()
    texels = {}
    for name in write_files(dst):
        shutil.copyfile(os.path.join(dst, name + ".tga"), os.path.join(tex_dir, f"tg_{name}.tga"))
        with open(os.path.join(harness.RUN_MODELS, f"tg_{name}.obj"), "w") as f:
            f.write(obj.replace("mtllib quadbox.mtl", f"mtllib tg_{name}.mtl"))
        with open(os.path.join(harness.RUN_MODELS, "materials", f"tg_{name}.mtl"), "w") as f:
            f.write(f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/tg_{name}.tga\n")
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        with open(txt, "w") as f:
            f.write(scenes.scene_text("cornellObj", width=16, height=16, obj_path=f"../models/tg_{name}.obj"))
        b2s = os.path.join(d, "s.b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
        ref = PodScene.load(b2s)
        assert len(ref.textures) == 1, (name, len(ref.textures))
        texels[name] = ref.textures[0]
        print(name, ref.textures[0].shape, os.path.getsize(os.path.join(dst, name + ".tga")))
        shutil.rmtree(d)
    np.savez_compressed(os.path.join(dst, "texels.npz"), **texels)


if __name__ == "__main__":
    main()

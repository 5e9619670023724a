"""PNG textures of every common layout, decoded by the REFERENCE'S OWN loader (scene.cpp -> stb_image,
oracle/_ref/ref_cpu --b2s).

PIL writes small files (at least 6168 bytes of texels each: scene.cpp:148 prints data[2055*3 ...]): 8-bit grey / grey+alpha / RGB / RGBA, palette images with and without a tRNS chunk,
2- and 4-bit palettes, a 1-bit image, 16-bit grey, RGB and grey with a transparent colour (tRNS), an
uncompressed one (stored deflate blocks) and a 220x160 noisy one (many dynamic-Huffman blocks, matches that
reach far back in the 32 KiB window).  Each becomes the map_Kd of tests/golden/quadbox.obj and goes through
the reference's loader.  png/<name>.png is the input, png/texels.npz[<name>] the texels the reference holds after
loading (H x W x C, rows flipped as scene.cpp:133 does).
Needs /root/reference (through oracle/_ref) and PIL; the outputs are committed.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import harness  # noqa: E402
from mygpuraytracer_b200 import scenes  # noqa: E402
from mygpuraytracer_b200.podscene import PodScene  # noqa: E402


def pictures():
    from PIL import Image

    rng = np.random.default_rng(0xB200)

    def rgb(w, h, noise=10):
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        img = np.stack([127 + 120 * np.sin(x * 0.31 + y * 0.17), 127 + 120 * np.cos(y * 0.23 - x * 0.11),
                        127 + 120 * np.sin((x + y) * 0.19)], -1) + rng.normal(0, noise, (h, w, 3))
        return np.clip(img, 0, 255).astype(np.uint8)

    out = {}
    out["rgb_93x71"] = (Image.fromarray(rgb(93, 71)), {})
    out["rgba_85x77"] = (Image.fromarray(np.dstack([rgb(85, 77), rng.integers(0, 256, (77, 85), dtype=np.uint8)]), "RGBA"), {})
    out["grey_89x83"] = (Image.fromarray(rgb(89, 83)[..., 0], "L"), {})
    out["greyalpha_86x79"] = (Image.fromarray(rgb(86, 79)[..., :2].copy(), "LA"), {})
    pal = Image.fromarray(rgb(90, 80)).quantize(colors=200)
    out["palette_90x80"] = (pal, {})
    out["palette_trns_90x80"] = (pal.copy(), {"transparency": bytes(rng.integers(0, 256, 120, dtype=np.uint8).tolist())})
    out["palette4bit_83x77"] = (Image.fromarray(rgb(83, 77)).quantize(colors=13), {"bits": 4})
    out["palette2bit_81x79"] = (Image.fromarray(rgb(81, 79)).quantize(colors=4), {"bits": 2})
    out["onebit_97x81"] = (Image.fromarray((rgb(97, 81)[..., 0] > 127).astype(np.uint8) * 255).convert("1"), {})
    out["grey16_92x80"] = (Image.fromarray((rng.integers(0, 65536, (80, 92))).astype(np.uint16)), {})
    t = rgb(88, 72, noise=0)
    t[30:60, 40:70] = (10, 200, 30)
    out["rgb_trns_88x72"] = (Image.fromarray(t), {"transparency": (10, 200, 30)})
    g = rgb(94, 78, noise=0)[..., 0].copy()
    g[20:40, 30:70] = 77
    out["grey_trns_94x78"] = (Image.fromarray(g, "L"), {"transparency": 77})
    out["stored_70x60"] = (Image.fromarray(rgb(70, 60)), {"compress_level": 0})
    out["noisy_220x160"] = (Image.fromarray(rgb(220, 160, noise=40)), {"compress_level": 9})
    smooth = np.tile(rgb(64, 8, noise=0), (25, 4, 1))  # long far-reaching matches
    out["repeats_256x200"] = (Image.fromarray(smooth), {"compress_level": 6})
    return out


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "png")
    os.makedirs(dst, exist_ok=True)
    tex_dir = os.path.join(os.path.dirname(harness.RUN_MODELS), "textures")
    os.makedirs(tex_dir, exist_ok=True)
    obj = open(os.path.join(HERE, "quadbox.obj")).read()
    texels = {}
    for name, (img, opts) in pictures().items():
        png = os.path.join(dst, name + ".png")
        img.save(png, "PNG", **opts)
        shutil.copyfile(png, os.path.join(tex_dir, f"pg_{name}.png"))
        with open(os.path.join(harness.RUN_MODELS, f"pg_{name}.obj"), "w") as f:
            f.write(obj.replace("mtllib quadbox.mtl", f"mtllib pg_{name}.mtl"))
        with open(os.path.join(harness.RUN_MODELS, "materials", f"pg_{name}.mtl"), "w") as f:
            f.write(f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/pg_{name}.png\n")
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        with open(txt, "w") as f:
            f.write(scenes.scene_text("cornellObj", width=16, height=16, obj_path=f"../models/pg_{name}.obj"))
        b2s = os.path.join(d, "s.b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
        ref = PodScene.load(b2s)
        assert len(ref.textures) == 1, len(ref.textures)
        texels[name] = ref.textures[0]
        print(name, ref.textures[0].shape, os.path.getsize(png))
        shutil.rmtree(d)
    np.savez_compressed(os.path.join(dst, "texels.npz"), **texels)


if __name__ == "__main__":
    main()

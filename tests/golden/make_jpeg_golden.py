"""Baseline JPEG files of several sampling layouts, decoded by the REFERENCE'S OWN loader (scene.cpp ->
stb_image, oracle/_ref/ref_cpu --b2s).

The seven textures the reference ships are 4:4:4; textures of other assets usually are not.  PIL writes small
baseline files with 4:2:0, 4:2:2, 4:4:4 sampling, odd sizes (partial MCUs, a chroma plane of width 1), restart
intervals, a greyscale one and progressive files of each kind; each becomes the map_Kd of tests/golden/quadbox.obj and goes through the
reference's loader.  jpeg/<name>.jpg is the input, jpeg/<name>.npy the texels the reference holds after loading
(H x W x C, rows flipped as scene.cpp:133 does).
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

# name -> (width, height, PIL save options, greyscale)
CASES = {
    "s420_64x48": (64, 48, dict(subsampling=2, quality=90), False),
    "s420_37x29": (37, 29, dict(subsampling=2, quality=85), False),
    "s422_50x20": (50, 20, dict(subsampling=1, quality=92), False),
    "s444_33x17": (33, 17, dict(subsampling=0, quality=95), False),
    "s420_2x2": (2, 2, dict(subsampling=2, quality=90), False),
    "s420_1x5": (1, 5, dict(subsampling=2, quality=90), False),
    "s420_restart_40x40": (40, 40, dict(subsampling=2, quality=80, restart_marker_blocks=2), False),
    "grey_21x13": (21, 13, dict(quality=90), True),
    # progressive files (spectral selection + successive approximation); large enough for scene.cpp:148
    "prog420_90x70": (90, 70, dict(subsampling=2, quality=88, progressive=True), False),
    "prog444_57x43": (57, 43, dict(subsampling=0, quality=93, progressive=True), False),
    "prog422_70x50": (70, 50, dict(subsampling=1, quality=75, progressive=True, optimize=True), False),
    "prog_grey_95x77": (95, 77, dict(quality=85, progressive=True), True),
    "prog420_restart_88x72": (88, 72, dict(subsampling=2, quality=80, progressive=True, restart_marker_blocks=3), False),
}


def picture(w, h, grey, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x * 0.9 + y * 0.3), 127 + 120 * np.cos(y * 1.1 - x * 0.2),
                    127 + 120 * np.sin((x + y) * 0.7)], -1) + rng.normal(0, 12, (h, w, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    return img[..., 0] if grey else img


def main():
    from PIL import Image

    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "jpeg")
    os.makedirs(dst, exist_ok=True)
    tex_dir = os.path.join(os.path.dirname(harness.RUN_MODELS), "textures")
    os.makedirs(tex_dir, exist_ok=True)
    obj = open(os.path.join(HERE, "quadbox.obj")).read()
    for k, (name, (w, h, opts, grey)) in enumerate(CASES.items()):
        jpg = os.path.join(dst, name + ".jpg")
        Image.fromarray(picture(w, h, grey, 100 + k)).save(jpg, "JPEG", **opts)
        shutil.copyfile(jpg, os.path.join(tex_dir, f"jg_{name}.jpg"))
        with open(os.path.join(harness.RUN_MODELS, f"jg_{name}.obj"), "w") as f:
            f.write(obj.replace("mtllib quadbox.mtl", f"mtllib jg_{name}.mtl"))
        with open(os.path.join(harness.RUN_MODELS, "materials", f"jg_{name}.mtl"), "w") as f:
            f.write(f"newmtl plain\nKd 0.5 0.5 0.5\nmap_Kd ../textures/jg_{name}.jpg\n")
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        with open(txt, "w") as f:
            f.write(scenes.scene_text("cornellObj", width=16, height=16, obj_path=f"../models/jg_{name}.obj"))
        b2s = os.path.join(d, "s.b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
        ref = PodScene.load(b2s)
        assert len(ref.textures) == 1, len(ref.textures)
        np.save(os.path.join(dst, name + ".npy"), ref.textures[0])
        print(name, ref.textures[0].shape, os.path.getsize(jpg))
        shutil.rmtree(d)


if __name__ == "__main__":
    main()

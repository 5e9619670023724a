"""MTL texture statements, loaded by the REFERENCE'S OWN scene.cpp + tinyobjloader + stb_image.

Which of the four texture slots of an OBJ geom get a texture, and which texels, for: options in front of the
file name (-s u v w, -bm f, -clamp on -blendu off, -o with a non-numeric argument that tinyobjloader swallows
anyway), the `bump` / `map_bump` / `map_Bump` spellings, a file name with blanks, all four maps at once, and a
map whose file does not exist (the reference keeps an empty texture and goes on).  The textures are PNG
fixtures of tests/golden/png.  mtl_textures/<name>.mtl is the input; mtl_textures/expected.npz holds, per
variant, the four slot indices of the geom and a CRC-32 of every texture the reference loaded.
Needs /root/reference (through oracle/_ref); the outputs are committed.
"""
import os
import shutil
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import harness  # noqa: E402
from mygpuraytracer_b200 import scenes  # noqa: E402
from mygpuraytracer_b200.podscene import PodScene  # noqa: E402

TEXTURES = {"opt tex a.png": "rgb_93x71.png", "opt_b.png": "rgba_85x77.png", "opt_c.png": "palette_90x80.png",
            "opt_d.png": "stored_70x60.png"}
HEAD = "newmtl plain\nKd 0.5 0.5 0.5\n"
MTLS = {
    "options_s": HEAD + "map_Kd -s 1 1 1 ../textures/opt_b.png\n",
    "bump_bm": HEAD + "map_Bump -bm 0.5 ../textures/opt_c.png\n",
    "bump_kw": HEAD + "bump ../textures/opt_c.png\n",
    "map_bump_lower": HEAD + "map_bump ../textures/opt_c.png\n",
    "all_four": HEAD + "map_Kd ../textures/opt_b.png\nmap_Ks ../textures/opt_c.png\nmap_Ke ../textures/opt_d.png\n"
                       "map_Bump ../textures/opt tex a.png\n",
    "clamp_blend": HEAD + "map_Kd -clamp on -blendu off ../textures/opt_d.png\n",
    "swallowed_name": HEAD + "map_Kd -o 0.5 0.5 ../textures/opt_d.png\nmap_Ks ../textures/opt_c.png\n",
    "mm_type": HEAD + "map_Ke -mm 0 1 -type sphere -imfchan r ../textures/opt_d.png\n",
    "missing_file": HEAD + "map_Kd ../textures/nope.png\nmap_Ks ../textures/opt_c.png\n",
    "name_with_blanks": HEAD + "map_Kd ../textures/opt tex a.png\n",
}
SLOTS = ("tex_kd", "tex_ks", "tex_bump", "tex_ke")


def summary(pod: PodScene):
    return (np.array([int(pod.geoms[k][6]) for k in SLOTS], np.int32),
            np.array([zlib.crc32(np.ascontiguousarray(t).tobytes()) for t in pod.textures], np.uint32))


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "mtl_textures")
    os.makedirs(dst, exist_ok=True)
    tex_dir = os.path.join(os.path.dirname(harness.RUN_MODELS), "textures")
    os.makedirs(tex_dir, exist_ok=True)
    for name, src in TEXTURES.items():
        shutil.copyfile(os.path.join(HERE, "png", src), os.path.join(tex_dir, name))
    obj = open(os.path.join(HERE, "quadbox.obj")).read()
    expected = {}
    for name, mtl in MTLS.items():
        with open(os.path.join(harness.RUN_MODELS, f"mo_{name}.obj"), "w") as f:
            f.write(obj.replace("mtllib quadbox.mtl", f"mtllib mo_{name}.mtl"))
        with open(os.path.join(harness.RUN_MODELS, "materials", f"mo_{name}.mtl"), "w") as f:
            f.write(mtl)
        with open(os.path.join(dst, name + ".mtl"), "w") as f:
            f.write(mtl)
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        with open(txt, "w") as f:
            f.write(scenes.scene_text("cornellObj", width=16, height=16, obj_path=f"../models/mo_{name}.obj"))
        b2s = os.path.join(d, "s.b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
        slots, crcs = summary(PodScene.load(b2s))
        expected[name + "_slots"], expected[name + "_crcs"] = slots, crcs
        print(name, slots.tolist(), [f"{c:08x}" for c in crcs])
        shutil.rmtree(d)
    np.savez(os.path.join(dst, "expected.npz"), **expected)


if __name__ == "__main__":
    main()

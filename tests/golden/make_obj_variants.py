"""OBJ statement corner cases, loaded by the REFERENCE'S OWN scene.cpp + tinyobjloader (oracle/_ref/ref_cpu).

Variants of tests/golden/quadbox.obj: vertex colours and a w coordinate after the position, a one-component
vt, line / point elements and free-form statements between the faces, trailing blanks, no newline at the end
of the file, an unknown usemtl, `mtllib` naming two files of which the first does not exist, faces written as
v/vt, and a "face" of two corners.  obj_variants/<name>.obj is the input, obj_variants/<name>.npz the faces
(positions, uv) and the appended material as the reference's loader produced them.
Needs /root/reference (through oracle/_ref); the outputs are committed.
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


def variants():
    base = open(os.path.join(HERE, "quadbox.obj")).read()
    return {
        "vertex_colors": base.replace("v 0.0 2.0 2.0", "v 0.0 2.0 2.0 1.0 0.5 0.25").replace("v 2.0 0.0 2.0", "v 2.0 0.0 2.0 0.1 0.2 0.3"),
        "v_w": base.replace("v 0.0 0.0 2.0", "v 0.0 0.0 2.0 1.0"),
        "vt_one": base.replace("vt 0.5 0.5", "vt 0.5"),
        "lines_points": base.replace("g roof", "l 1 2 3\np 4\ng roof"),
        "trailing_ws": base.replace("f 5/1/1 1/2/1 9/5/1", "f 5/1/1 1/2/1 9/5/1   \t"),
        "no_newline_end": base.rstrip("\n"),
        "vp_and_unknown": base.replace("vn 0.0 0.0 1.0", "vn 0.0 0.0 1.0\nvp 0.1 0.2 0.3\ncstype bezier\ndeg 3"),
        "usemtl_unknown": base.replace("usemtl plain", "usemtl doesnotexist"),
        "two_mtllibs": base.replace("mtllib quadbox.mtl", "mtllib nothere.mtl quadbox.mtl"),
        "face_v_vt": base.replace("f 5/1/1 1/2/1 9/5/1", "f 5/1 1/2 9/5"),
        "two_corner_face": base.replace("g roof", "f 1/1/1 2/2/1\ng roof"),
    }


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "obj_variants")
    os.makedirs(dst, exist_ok=True)
    shutil.copyfile(os.path.join(HERE, "quadbox.mtl"), os.path.join(harness.RUN_MODELS, "materials", "quadbox.mtl"))
    for name, obj in variants().items():
        with open(os.path.join(harness.RUN_MODELS, f"ov_{name}.obj"), "w") as f:
            f.write(obj)
        with open(os.path.join(dst, name + ".obj"), "w") as f:
            f.write(obj)
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        with open(txt, "w") as f:
            f.write(scenes.scene_text("cornellObj", width=16, height=16, obj_path=f"../models/ov_{name}.obj"))
        b2s = os.path.join(d, "s.b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
        ref = PodScene.load(b2s)
        np.savez_compressed(os.path.join(dst, name + ".npz"), face_pos=ref.face_pos, face_uv=ref.face_uv,
                            material=np.frombuffer(ref.materials[-1:].tobytes(), np.uint8))
        print(name, len(ref.face_pos))
        shutil.rmtree(d)


if __name__ == "__main__":
    main()

"""MTL corner cases, loaded by the REFERENCE'S OWN scene.cpp + tinyobjloader (oracle/_ref/ref_cpu --b2s).

The reference appends ONE material per OBJ geom, filled from the first `newmtl` of the MTL file
(apps/src/scene.cpp:68,134,220-231).  Each variant is the MTL of tests/golden/quadbox.obj rewritten to exercise
one property: an empty material (tinyobjloader's defaults), missing Ni, Ke (becomes the emittance), tabs / CRLF /
trailing blanks, comments and statements the loader ignores, and two materials of which only the first counts.
mtl_variants/<name>.mtl is the input, <name>_material.npy the 44 bytes of the material the reference appended.
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

MTLS = {
    "minimal": "newmtl a\n",
    "no_ni": "newmtl a\nKd 0.1 0.2 0.3\nKs 0.5 0.6 0.7\n",
    "ke": "newmtl a\nKd 0.1 0.2 0.3\nKe 0.4 0.5 0.6\nNi 1.7\nNs 55\n",
    "tabs_crlf": "newmtl a\r\n\tKd\t0.25 0.5 0.75\r\n  Ks 1 1 1  \r\nNi\t2\r\n",
    "comments_unknown": "# c\nnewmtl a\nKd 0.3 0.3 0.3\nfoo 1 2 3\nTf 1 1 1\nd 0.5\nillum 7\nNi 1.1\n",
    "second_first": "newmtl z\nKd 0.9 0.8 0.7\nNi 3\nnewmtl a\nKd 0.1 0.1 0.1\n",
}


def obj_for(name: str) -> str:
    src = open(os.path.join(HERE, "quadbox.obj")).read()
    return src.replace("mtllib quadbox.mtl", f"mtllib mv_{name}.mtl").replace("usemtl plain", "usemtl a")


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "mtl_variants")
    os.makedirs(dst, exist_ok=True)
    for name, mtl in MTLS.items():
        with open(os.path.join(harness.RUN_MODELS, "materials", f"mv_{name}.mtl"), "wb") as f:
            f.write(mtl.encode())
        with open(os.path.join(harness.RUN_MODELS, f"mv_{name}.obj"), "w") as f:
            f.write(obj_for(name))
        d = harness.tmpdir()
        txt = os.path.join(d, "s.txt")
        with open(txt, "w") as f:
            f.write(scenes.scene_text("cornellObj", width=16, height=16, obj_path=f"../models/mv_{name}.obj"))
        b2s = os.path.join(d, "s.b2s")
        harness.run("ref_cpu", txt, os.path.join(d, "out"), b2s, iters=1, dump_iter=1)
        ref = PodScene.load(b2s)
        with open(os.path.join(dst, name + ".mtl"), "wb") as f:
            f.write(mtl.encode())
        np.save(os.path.join(dst, name + "_material.npy"), np.frombuffer(ref.materials[-1:].tobytes(), np.uint8))
        print(name, ref.materials[-1])
        shutil.rmtree(d)


if __name__ == "__main__":
    main()

"""Scene-file format corner cases, loaded by the REFERENCE'S OWN scene.cpp (oracle/_ref/ref_cpu --b2s).

Each variant of cornellGlass.txt exercises one property of the text format of apps/src/scene.cpp
(SURVEY.md 8b): camera orbit recompute from other EYE / LOOKAT / UP / FOVY values and a non-square RES,
tabs and number spellings (.98, 9.8e-1, +.98), TRANS / ROTAT / SCALE lines in another order, a file without
a trailing newline, a `triangle` geom, non-trivial rotations with a non-uniform scale, a material that is
reflective and refractive at once, other ITERATIONS / DEPTH.  variants/<name>.txt is the input,
variants/<name>.b2s the scene as the reference's loader produced it.
Needs /root/reference (through oracle/_ref); the outputs are committed.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import harness  # noqa: E402
from mygpuraytracer_b200 import scenes  # noqa: E402


def variants():
    base = scenes.scene_text("cornellGlass", width=48, height=20)
    out = {}
    out["camera"] = (base.replace("FOVY        45", "FOVY        30.5").replace("EYE         0.0 5 10.5", "EYE         2.5 3 9.25")
                     .replace("LOOKAT      0 5 0", "LOOKAT      -1 4.5 0.5").replace("UP          0 1 0", "UP          0.1 1 0"))
    out["tabs_numbers"] = (base.replace("RGB         .98 .98 .98", "RGB\t9.8e-1\t+.98 0.98")
                           .replace("TRANS       0 10 0", "TRANS  \t 0 1e1 0"))
    v = base.split("\n")
    i = v.index("OBJECT 3")
    v[i + 3], v[i + 4], v[i + 5] = v[i + 5], v[i + 3], v[i + 4]
    out["permuted_transform"] = "\n".join(v)
    out["no_trailing_newline"] = base.rstrip("\n")
    out["triangle_geom"] = base + "\nOBJECT 7\ntriangle\nmaterial 2\nTRANS       0 2 0\nROTAT       10 20 30\nSCALE       1 2 3\n"
    out["rot_scale"] = (base.replace("ROTAT       0 0 90", "ROTAT       33 -47 91.5")
                        .replace("SCALE       3 3 3", "SCALE       1.5 2.25 0.75"))
    out["both_refl_refr"] = base.replace("REFL        0\nREFR        1", "REFL        1\nREFR        1")
    out["iter_depth"] = base.replace("ITERATIONS  5000", "ITERATIONS  17").replace("DEPTH       8", "DEPTH       3")
    for name, txt in out.items():
        assert txt != base or name == "no_trailing_newline", name
    return out


def main():
    assert harness.have("ref_cpu"), "build oracle/_ref first: make -C oracle ref"
    dst = os.path.join(HERE, "variants")
    os.makedirs(dst, exist_ok=True)
    for name, txt in variants().items():
        p = os.path.join(dst, name + ".txt")
        with open(p, "w") as f:
            f.write(txt)
        d = harness.tmpdir()
        harness.run("ref_cpu", p, os.path.join(d, "out"), os.path.join(dst, name + ".b2s"), iters=1, dump_iter=1)
        shutil.rmtree(d)
        print(name, os.path.getsize(os.path.join(dst, name + ".b2s")))


if __name__ == "__main__":
    main()

"""Build experimental variants of libb2pt.so: `name=-DFLAG1,-DFLAG2 ...` -> build/variants/libb2pt_<name>.so.

Select one at run time with B2PT_LIB=<path> (see mygpuraytracer_b200/api.py).
"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import build as B

out_dir = os.path.join(os.path.dirname(B.HERE), "build", "variants")
os.makedirs(out_dir, exist_ok=True)

def one(spec):
    name, _, flags = spec.partition("=")
    out = os.path.join(out_dir, f"libb2pt_{name}.so")
    cmd = ["nvcc"] + B.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-Xptxas", "-v"] + B.sources() + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return f"{name}: FAILED\n{r.stderr[-2000:]}"
    info = []
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):
        if "Compiling entry function" in l and ("k_mesh_walk" in l or "k_intersect_analytic" in l):
            info.append(l.split("'")[1][:40] + " :: " + lines[i + 2].strip() + " | " + lines[i + 3].strip()[:60])
    return f"{name}: ok\n  " + "\n  ".join(info)

with ThreadPoolExecutor(4) as ex:
    for res in ex.map(one, sys.argv[1:]):
        print(res, flush=True)

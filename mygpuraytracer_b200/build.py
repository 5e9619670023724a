"""Build ``libb2pt.so`` in-tree (nvcc, sm_100a).

The shared object holds the hand-written CUDA kernels (``csrc/*.cuh``), the
context / C ABI (``csrc/b2pt.cu``) and the host-side scene loader
(``csrc/host``).  Flags that matter:

* ``-gencode arch=compute_100a,code=sm_100a``: Blackwell B200 only, no PTX for
  other architectures, no multi-backend dispatch;
* ``-fmad=false``: no FMA contraction, so the kernels round exactly as the
  reference's expressions are written (see ``csrc/pt_math.cuh``);
* ``-lineinfo``: ncu source pages map to these files;
* ``-fwrapv`` (host code): the integer IDCT of the JPEG decoder overflows on damaged files exactly as
  stb_image's does; wrapping makes that defined behaviour.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2pt.so")
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-fmad=false", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fwrapv,-Wall",
    "--shared",
]


def sources():
    return [os.path.join(CSRC, "b2pt.cu"), os.path.join(CSRC, "pipe.cu"), os.path.join(CSRC, "multi.cu"), os.path.join(CSRC, "host", "scene_loader.cpp")]


def _deps():
    out = []
    for root, _d, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(HERE, "..", "include", "b2pt.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

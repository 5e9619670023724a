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


# The same library with 8-wide BVH nodes (two 128-byte lines per node, -DB2PT_WIDE=8): built and kept beside the
# product so that tests/test_gpu_wide8.py can run the parity cases through it.  It is NOT what the renderer loads:
# measured on the B200 it is slower than the 4-wide walk (walk 0.497 -> 0.644 ms per iteration for 26 % fewer
# steps and a third fewer hand-offs: profiles/r02_notes.md).
LIB_WIDE8 = os.path.join(HERE, "libb2pt_wide8.so")


def build_wide8(force: bool = False) -> str:
    if not force and os.path.exists(LIB_WIDE8) and all(os.path.getmtime(p) <= os.path.getmtime(LIB_WIDE8) for p in _deps()):
        return LIB_WIDE8
    r = subprocess.run(["nvcc"] + NVCC_FLAGS + ["-DB2PT_WIDE=8"] + sources() + ["-o", LIB_WIDE8], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed (wide8):\n" + r.stdout + r.stderr)
    return LIB_WIDE8


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--wide8" in sys.argv:
        print(build_wide8(force="--force" in sys.argv))

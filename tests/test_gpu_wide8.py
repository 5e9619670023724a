"""The walk with 8-wide BVH nodes (-DB2PT_WIDE=8, mygpuraytracer_b200/libb2pt_wide8.so: two 128-byte lines per node,
nearest child first and the other hits unsorted, 20 stack entries per lane): built by build() beside the product
library and run here through the same parity cases -- LBVH walk against the oracle's brute-force loop, one and two
meshes, forced hand-offs -- in a child process, because a process loads one build of the library.  The variant is
measured and NOT shipped as the product path (profiles/r02_notes.md): it takes 26 % fewer steps and hands a third
fewer walks off, and is slower."""
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu
LIB = os.path.join(ROOT, "mygpuraytracer_b200", "libb2pt_wide8.so")


def test_wide8_walk_is_bit_exact():
    if not os.path.exists(LIB):
        pytest.skip("libb2pt_wide8.so has not been built (python -m mygpuraytracer_b200.build --wide8)")
    env = dict(os.environ, B2PT_LIB=LIB)
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-k", "mesh_scenes_bvh_bitexact or two_meshes or bvh_equals_brute_force"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert " passed" in p.stdout and "failed" not in p.stdout, p.stdout[-2000:]

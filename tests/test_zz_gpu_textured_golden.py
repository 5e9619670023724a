"""The textured-mesh golden (tests/golden/texquad_32x32: a 4-channel kd map, a palette ks map, a bump map and an
emission map on tests/golden/texquad.obj, dumped by the reference's own host build) through the CUDA path.

Kept in a file that sorts last: it was added after the round's last GPU session, so it has run on the CPU side
(oracle == reference, tests/test_oracle_golden.py) but not yet on a GPU; under `pytest -x` a surprise here cannot
hide any other test.
"""
import os

import pytest

from mygpuraytracer_b200.podscene import PodScene
from test_gpu_parity import compare_iteration
from util import GOLDEN

pytestmark = pytest.mark.gpu


def test_textured_mesh_golden_every_stage_bitexact():
    pod = PodScene.load(os.path.join(GOLDEN, "texquad_32x32.b2s"))
    assert [t.shape[2] for t in pod.textures] == [4, 3, 3, 3]
    compare_iteration(pod, {}, what="texquad_32x32")


def test_grey_and_grey_alpha_maps_every_stage_bitexact():
    """tests/golden/greyquad_32x32: a 1-channel kd map and a 2-channel ks map.  The reference reads three bytes per
    texel whatever the channel count (apps/src/interactions.h:199-212), i.e. into the following texels; the scene and
    every stage come from the reference's own host build."""
    pod = PodScene.load(os.path.join(GOLDEN, "greyquad_32x32.b2s"))
    assert sorted(t.shape[2] for t in pod.textures) == [1, 2, 3, 3]
    compare_iteration(pod, {}, what="greyquad_32x32")

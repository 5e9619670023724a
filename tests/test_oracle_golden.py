"""The oracle against the known-answer vectors of SURVEY.md section 8c
(generated from the reference's own headers) and against the committed stage
dumps of the reference's host build (tests/golden, made by make_golden.py)."""
import ctypes as C

import numpy as np
import pytest

from mygpuraytracer_b200 import abi
from oracle import oracle
from util import STAGE_MAP, assert_same_bits, load_golden


def f32(x):
    return np.float32(x)


def test_minstd_10000th_value():
    # thrust::minstd_rand default-seeded, documented 10000th value
    assert oracle.minstd_nth(10000, 1) == 399268537


def test_seed_and_draws():
    # makeSeededRandomEngine(1, 12345, 8): h = 939298829, draws in (-0.5, 0.5)
    assert oracle.seed(1, 12345, 8) == 939298829 % 2147483647
    d = oracle.draws(1, 12345, 8, 2, -0.5, 0.5)
    assert f32(d[0]) == f32(-0.00102737546) and f32(d[1]) == f32(0.407480478)
    d = oracle.draws(1, 0, 0, 3)
    assert [f32(x) for x in d] == [f32(0.00288156001), f32(0.0958057344), f32(0.638674617)]
    d = oracle.draws(7, 2073599, 0, 2)
    assert [f32(x) for x in d] == [f32(0.0678336099), f32(0.396139026)]


def test_hemisphere_vector():
    v = oracle.hemisphere([0, 1, 0], 1, 0, 0)
    assert_same_bits(v, np.array([-0.565446854, 0.0536801629, -0.823036015], np.float32), "hemisphere")


def _geom(scene, idx):
    g = abi.Geom()
    C.memmove(C.byref(g), scene.geoms[idx:idx + 1].ctypes.data, C.sizeof(abi.Geom))
    return g


def test_sphere_and_box_vectors():
    scene, *_ = load_golden("cornell_32x32")
    o = np.array([0, 5, 10.5], np.float32)
    # sphere TRANS -1 4 -1 SCALE 3: ray toward its centre
    d = np.array([-1, 4, -1], np.float32) - o
    d = (d / np.float32(np.sqrt(np.float32((d * d).sum())))).astype(np.float32)
    t, n = oracle.geom_test("sphere", _geom(scene, 6), o, d)
    assert abs(t - 10.0863266) < 2e-6
    assert np.allclose(n, [0.0863064229, 0.0863064826, 0.992523253], atol=2e-7)
    # back wall cube, dir normalize(0.3, -0.2, -1)
    d = np.array([0.3, -0.2, -1.0], np.float32)
    d = (d * (np.float32(1) / np.sqrt(np.float32((d * d).sum())))).astype(np.float32)
    t, n = oracle.geom_test("box", _geom(scene, 3), o, d)
    assert f32(t) == f32(16.4714108)
    assert_same_bits(n, np.array([4.37113847e-08, 0, 0.99999994], np.float32), "box normal")


def test_ray_triangle_vector():
    hit, bary = oracle.ray_triangle([.25, .25, 1], [0, 0, -1], [0, 0, 0], [1, 0, 0], [0, 1, 0])
    assert hit and list(bary) == [0.25, 0.25, 1.0]
    hit, _ = oracle.ray_triangle([.25, .25, -1], [0, 0, 1], [0, 0, 0], [1, 0, 0], [0, 1, 0])
    assert not hit  # back face culled


def test_portable_sincos_accuracy():
    x = np.linspace(-2.0, 7.0, 20001).astype(np.float32)
    s, c = oracle.sincos_portable(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 2.5e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 2.5e-7


CASES = [
    ("cornell_32x32", {}), ("cornellGlass_32x32", {}), ("cornellGlass_dof_32x24", {"depth_of_field": 1}),
    ("cornellGlass_noaa_24x32", {"antialiasing": 0}), ("sphere_16x16", {}), ("quadbox_32x32", {}),
    ("hardobj_16x16", {}), ("rot_scale_48x20", {}), ("both_refl_refr_48x20", {}),
    ("texquad_32x32", {}), ("greyquad_32x32", {}),
]


@pytest.mark.parametrize("case,optkw", CASES)
def test_oracle_matches_reference_host_build(case, optkw):
    """Every stage of iteration 1 and the image after two iterations are
    bit-identical to the reference's own code (libm trig on both sides)."""
    scene, ref_depths, ref_image, ref_albedo, ref_nlive = load_golden(case)
    opt = abi.default_options(**optkw)
    oracle.set_dof_arg_order("dof" in case)  # fixtures come from the g++ host build
    try:
        image = np.zeros((scene.n_pixels, 3), np.float32)
        albedo = np.zeros((scene.n_pixels, 3), np.float32)
        st = oracle.iteration_with_stages(scene, opt, 1, image, albedo)
        assert len(st) == len(ref_depths)
        for d, (a, b) in enumerate(zip(st, ref_depths)):
            for x, y in STAGE_MAP:
                assert_same_bits(a[x], b[y], f"{case} depth {d} {x}")
            hit = a["hit_t"] > 0
            assert np.array_equal(a["hit_geom"][hit], b["hit_geom"][hit])
            assert int(ref_nlive[d]) == len(a["ray_pixel"])
        oracle.iteration_with_stages(scene, opt, 2, image, None)
        assert_same_bits(image, ref_image, "image after 2 iterations")
        assert_same_bits(albedo, ref_albedo, "albedo")
        # the all-in-C loop agrees with the stage-by-stage drive
        img2, alb2, _nl, seg = oracle.render(scene, opt, 1, 2, 1)
        assert_same_bits(img2, ref_image, "oracle_render image")
        assert_same_bits(alb2, ref_albedo, "oracle_render albedo")
        assert seg > 0
    finally:
        oracle.set_dof_arg_order(False)


def test_sort_and_partition_perms_are_stable():
    rng = np.random.default_rng(5)
    m = rng.integers(0, 7, 5000).astype(np.int32)
    perm = oracle.sort_perm(m)
    assert np.array_equal(perm, np.argsort(-m.astype(np.int64), kind="stable"))
    b = rng.integers(0, 3, 5000).astype(np.int32)
    pperm, live = oracle.partition_perm(b)
    idx = np.arange(len(b))
    assert live == int((b > 0).sum())
    assert np.array_equal(pperm, np.concatenate([idx[b > 0], idx[b <= 0]]))


def test_empty_and_tiny_inputs():
    assert len(oracle.sort_perm(np.zeros(0, np.int32))) == 0
    p, live = oracle.partition_perm(np.zeros(0, np.int32))
    assert live == 0 and len(p) == 0
    p, live = oracle.partition_perm(np.array([0], np.int32))
    assert live == 0 and list(p) == [0]

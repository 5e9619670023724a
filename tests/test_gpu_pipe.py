"""b2pt_pipe_*: pathtrace() one iteration per call with the next iterations rendered ahead (csrc/pipe.cu).

The contract is the reference's (apps/src/main.cpp:255, apps/src/pathtrace.cu:663-668): after call i the host
image is the running sum including iteration i.  Every case compares the pipeline with b2pt_pathtrace on a
single context, call by call and BIT FOR BIT: a lane's contribution to a pixel is 0 or color*PI, so merging
the lanes in call order reproduces the sequential `image[pixel] += color * PI`.
"""
import numpy as np
import pytest

from mygpuraytracer_b200 import abi, api, assets, scenes
from util import assert_same_bits

pytestmark = pytest.mark.gpu


def _mesh_scene(tmp_path, name, w, h, tris):
    root = assets.prepare(str(tmp_path / "run"), triangles=tris, procedural_size=256)
    return api.Scene(assets.scene_file(name, w, h, root=root)).pod


def _sequence(pod, iters, lanes, expect_misses=None):
    n = pod.n_pixels
    img_a, alb_a = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    img_b, alb_b = np.full((n, 3), -1, np.float32), np.full((n, 3), -1, np.float32)
    with api.Renderer(pod, abi.default_options()) as one, api.Pipeline(pod, abi.default_options(), lanes=lanes) as pipe:
        for k, it in enumerate(iters):
            one.pathtrace(it, img_a, alb_a)
            pipe.pathtrace(it, img_b, alb_b)
            assert_same_bits(img_a, img_b, f"running sum after call {k} (iteration {it}, {lanes} lanes)")
            assert_same_bits(alb_a, alb_b, f"albedo after call {k}")
        assert img_a.any(), "the scene must produce light"
        if expect_misses is not None:
            assert pipe.misses() == expect_misses
        assert pipe.launch_count() > 0


@pytest.mark.parametrize("lanes", [1, 2, 4, 7])
def test_consecutive_iterations_bitexact(tmp_path, lanes):
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 128, 72, 5000)
    _sequence(pod, list(range(1, 14)), lanes, expect_misses=0)


def test_analytic_scene_and_odd_size(tmp_path):
    # 131*77*3 floats is not a multiple of 4: the merge kernel's scalar tail
    pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=131, height=77)).pod
    _sequence(pod, list(range(1, 10)), 4, expect_misses=0)


def test_strided_sequence_is_learned(tmp_path):
    """A rank of a samples-per-pixel sharded job calls r+1, r+1+R, ...: one miss, then the stride is predicted."""
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 96, 54, 1000)
    _sequence(pod, [2, 5, 8, 11, 14, 17, 20], 4, expect_misses=1)


def test_restart_and_irregular_calls(tmp_path):
    """Calls that do not continue the sequence drop the speculated iterations; the sum still holds every call."""
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 96, 54, 1000)
    _sequence(pod, [1, 2, 3, 1, 2, 9, 4, 5, 6, 7, 7, 8], 3)


def test_reset_and_camera_change(tmp_path):
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 96, 54, 1000)
    n = pod.n_pixels
    img_a, alb_a = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    img_b, alb_b = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    cam = pod.camera.copy()
    moved = cam.copy()
    moved["position"][0, 0] += 0.75
    with api.Renderer(pod, abi.default_options()) as one, api.Pipeline(pod, abi.default_options(), lanes=4) as pipe:
        for it in (1, 2, 3):
            one.pathtrace(it, img_a, alb_a)
            pipe.pathtrace(it, img_b, alb_b)
        one.reset()
        pipe.reset()
        for it in (1, 2):
            one.pathtrace(it, img_a, alb_a)
            pipe.pathtrace(it, img_b, alb_b)
            assert_same_bits(img_a, img_b, f"after reset, iteration {it}")
            assert_same_bits(alb_a, alb_b, f"albedo after reset, iteration {it}")
        one.set_camera(moved)
        pipe.reset(moved)
        for it in (1, 2, 3):
            one.pathtrace(it, img_a, alb_a)
            pipe.pathtrace(it, img_b, alb_b)
            assert_same_bits(img_a, img_b, f"after the camera moved, iteration {it}")
            assert_same_bits(alb_a, alb_b, f"albedo after the camera moved, iteration {it}")


def test_free_functions_with_pipeline_lanes(tmp_path):
    """pathtraceInit / pathtrace / pathtraceFree of the host mirror served by a pipeline."""
    path = scenes.write_scene("cornell", str(tmp_path / "s.txt"), width=64, height=64)
    ref_scene, scene = api.Scene(path), api.Scene(path)
    api.set_pipeline_lanes(1)
    api.pathtraceInit(ref_scene)
    for it in range(1, 6):
        api.pathtrace(None, 0, it)
    api.pathtraceFree()
    try:
        api.set_pipeline_lanes(4)
        api.pathtraceInit(scene)
        for it in range(1, 6):
            api.pathtrace(None, 0, it)
        api.pathtraceFree()
    finally:
        api.set_pipeline_lanes(1)
    assert_same_bits(ref_scene.state.image, scene.state.image, "state.image")
    assert_same_bits(ref_scene.state.albedo, scene.state.albedo, "state.albedo")


def test_pipe_rejects_bad_arguments(tmp_path):
    pod = api.Scene(scenes.write_scene("cornell", str(tmp_path / "s.txt"), width=32, height=32)).pod
    with pytest.raises(api.B2ptError):
        api.Pipeline(pod, abi.default_options(), lanes=0)
    with pytest.raises(api.B2ptError):
        api.Pipeline(pod, abi.default_options(record_stages=1), lanes=2)


def test_shared_scene_contexts_bitexact_and_outlive_their_parent(tmp_path):
    """b2pt_create_shared: contexts that use the parent's device copy of the scene render the same bits, also
    after the parent has been destroyed (the shared data lives until its last user goes)."""
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 128, 72, 5000)
    n = pod.n_pixels
    with api.Renderer(pod, abi.default_options()) as alone:
        alone.render(1, 3, 1)
        ref, ref_alb = alone.read()
    parent = api.Renderer(pod, abi.default_options(concurrent_contexts=3))
    kids = [api.Renderer(pod, abi.default_options(concurrent_contexts=3), share=parent) for _ in range(2)]
    assert kids[0].bvh_info(int(np.nonzero(pod.geoms["type"] == abi.OBJ)[0][0])).n_faces == parent.bvh_info(
        int(np.nonzero(pod.geoms["type"] == abi.OBJ)[0][0])).n_faces
    parent.render(1, 3, 1)
    img, alb = parent.read()
    assert_same_bits(ref, img, "parent of shared contexts")
    parent.close()
    for k in kids:  # both at once, on their own streams, after the parent is gone
        k.render(1, 3, 1)
    for k in kids:
        img, alb = k.read()
        assert_same_bits(ref, img, "shared context image")
        assert_same_bits(ref_alb, alb, "shared context albedo")
        k.close()
    assert n > 0

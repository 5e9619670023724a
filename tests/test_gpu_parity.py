"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle.

Integer / index work -- closest-hit geom and face ids, the material-sort
permutation, the compaction permutation, live counts -- must be bit-exact.
Floating point: with trig_mode=PORTABLE both sides evaluate the same operation
sequence without FMA contraction, so hit distances, normals, scattered rays,
per-iteration radiance and the accumulated image are compared BITWISE as well
(tolerance 0; north_star allows 1e-4 relative).  With trig_mode=NATIVE the
kernels call CUDA's sinf/cosf like the reference's device code; everything up
to the first diffuse bounce is still bit-exact against the oracle and the
converged image must reach PSNR >= 50 dB (north_star).
"""
import os

import numpy as np
import pytest

from mygpuraytracer_b200 import abi, api, assets, scenes
from mygpuraytracer_b200.podscene import PodScene
from oracle import oracle
from util import GOLDEN, assert_same_bits, load_golden, psnr

pytestmark = pytest.mark.gpu

STAGES_TO_COMPARE = ["ray_origin", "ray_dir", "ray_pixel", "hit_t", "hit_normal", "hit_material", "sort_perm",
                     "shaded_color", "shaded_bounces", "shaded_origin", "shaded_dir", "partition_pixel"]


def compare_iteration(pod: PodScene, optkw: dict, iters=(1, 2), what=""):
    opt = abi.default_options(trig_mode=abi.TRIG_PORTABLE, record_stages=1, **optkw)
    ref_img = np.zeros((pod.n_pixels, 3), np.float32)
    ref_alb = np.zeros((pod.n_pixels, 3), np.float32)
    with api.Renderer(pod, opt) as r:
        for it in iters:
            r.render(it, 1, 1)
            got = r.stages()
            ref = oracle.iteration_with_stages(pod, opt, it, ref_img, ref_alb if it == 1 else None)
            assert len(got) == len(ref), f"{what}: depth count {len(got)} vs {len(ref)}"
            for d, (g, o) in enumerate(zip(got, ref)):
                for name in STAGES_TO_COMPARE:
                    assert_same_bits(g[name], o[name], f"{what} iter {it} depth {d} {name}")
                hit = o["hit_t"] > 0
                assert np.array_equal(g["hit_geom"][hit], o["hit_geom"][hit]), f"{what} depth {d} geom ids"
                assert np.array_equal(g["hit_geom"][~hit], np.full((~hit).sum(), -1)), "miss geom id"
                assert np.array_equal(g["hit_face"][hit], o["hit_face"][hit]), f"{what} depth {d} face ids"
                is_obj = hit & (pod.geoms["type"][np.maximum(o["hit_geom"], 0)] == abi.OBJ)
                assert_same_bits(g["hit_uv"][is_obj], o["hit_uv"][is_obj], f"{what} depth {d} uv")
            live = r.live_counts()
            assert list(live[: len(ref)]) == [len(s["ray_pixel"]) for s in ref]
        img, alb = r.read()
    assert_same_bits(img, ref_img, f"{what} accumulated image")
    assert_same_bits(alb, ref_alb, f"{what} albedo")


@pytest.mark.parametrize("case,optkw", [
    ("cornell_32x32", {}), ("cornellGlass_32x32", {}), ("cornellGlass_dof_32x24", {"depth_of_field": 1}),
    ("cornellGlass_noaa_24x32", {"antialiasing": 0}), ("sphere_16x16", {}), ("quadbox_32x32", {}),
    ("hardobj_16x16", {}), ("rot_scale_48x20", {}), ("both_refl_refr_48x20", {}),
    ("cornell_32x32", {"sort_by_material": 0}),
])
def test_golden_scenes_every_stage_bitexact(case, optkw):
    pod = PodScene.load(os.path.join(GOLDEN, case + ".b2s"))
    compare_iteration(pod, optkw, what=case)


def test_reference_golden_dumps_direct():
    """Depth-0 hits and the first sort/partition against the REFERENCE'S OWN
    dump (tests/golden, libm trig): no trig has happened yet at that point, so
    the CUDA path must reproduce the reference bit for bit."""
    for case in ["cornell_32x32", "cornellGlass_32x32", "quadbox_32x32"]:
        pod, ref_depths, *_ = load_golden(case)
        with api.Renderer(pod, abi.default_options(record_stages=1)) as r:
            r.render(1, 1, 1)
            g = r.stages()[0]
        ref = ref_depths[0]
        for mine, theirs in [("ray_origin", "in_origin"), ("ray_dir", "in_dir"), ("hit_t", "hit_t"),
                             ("hit_normal", "hit_normal"), ("hit_material", "hit_material")]:
            assert_same_bits(g[mine], ref[theirs], f"{case} {mine}")
        assert np.array_equal(g["ray_pixel"][g["sort_perm"]], ref["sorted_pixel"]), "sort permutation"
        hit = ref["hit_t"] > 0
        assert np.array_equal(g["hit_geom"][hit], ref["hit_geom"][hit])


@pytest.mark.parametrize("name,w,h", [("cornell", 200, 150), ("cornellGlass", 160, 160)])
def test_larger_frames_bitexact(tmp_path, name, w, h):
    pod = api.Scene(scenes.write_scene(name, str(tmp_path / "s.txt"), width=w, height=h)).pod
    compare_iteration(pod, {}, iters=(1, 2, 7), what=f"{name} {w}x{h}")


@pytest.mark.parametrize("name,optkw", [("cornell", {}), ("cornellGlass", {"depth_of_field": 1, "antialiasing": 1})])
def test_baseline_configs_1_and_2_at_full_size(tmp_path, name, optkw):
    """BASELINE.json configs[0] and configs[1] at their own size (800x800, depth 8; cornellGlass with DOF and
    stochastic AA): image, albedo and live paths per depth after three iterations, bit for bit against the oracle's
    whole-iteration path (640 000 paths per iteration: 313 sort tiles, the analytic rounds with every geom type)."""
    pod = api.Scene(scenes.write_scene(name, str(tmp_path / "s.txt"), width=800, height=800)).pod
    opt = abi.default_options(trig_mode=abi.TRIG_PORTABLE, **optkw)
    with api.Renderer(pod, opt) as r:
        r.render(1, 3, 1)
        img, alb = r.read()
        live = r.live_counts()
    ref_img, ref_alb, ref_live, _seg = oracle.render(pod, opt, 1, 3, 1)
    assert_same_bits(img, ref_img, f"{name} 800x800 image after 3 iterations")
    assert_same_bits(alb, ref_alb, f"{name} 800x800 albedo")
    assert list(live[: len(ref_live)]) == list(ref_live[: len(live)])


def _mesh_scene(tmp_path, name, w, h, tris):
    root = assets.prepare(str(tmp_path / "run"), triangles=tris, procedural_size=256)
    return api.Scene(assets.scene_file(name, w, h, root=root)).pod


@pytest.mark.parametrize("name,tris", [("cornellObj", 1000), ("cornellSpaceship", 1000), ("cornellSpaceship", 20000)])
def test_mesh_scenes_bvh_bitexact(tmp_path, name, tris):
    """LBVH traversal + textures + bump map: closest-hit ids, uv, normals and
    everything downstream identical to the oracle's brute-force loop."""
    pod = _mesh_scene(tmp_path, name, 96, 54, tris)
    compare_iteration(pod, {}, what=f"{name}/{tris}")
    d0 = oracle.intersect(pod, *[a for a in (oracle.generate(pod, abi.default_options(), 1).origin,
                                             oracle.generate(pod, abi.default_options(), 1).dir)])
    assert (pod.geoms["type"][np.maximum(d0.geom, 0)][d0.t > 0] == abi.OBJ).sum() > 20, "mesh must be visible"


@pytest.mark.parametrize("env", [{"B2PT_SORT_GENERAL": "1"}, {"B2PT_SORT_GENERAL": "0"}])
def test_both_material_sorts_in_the_renderer(tmp_path, monkeypatch, env):
    """Scenes with at most 8 materials sort with k_sort_material_few, the others with the 256-bin k_sort_material:
    the same mesh scene through both (the switch only exists for this test and for A/B timing), every stage --
    the sort permutation and the compaction ranks behind partition_pixel included -- identical to the oracle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 173, 99, 5000)  # 17 127 paths: five sort tiles, the last one ragged
    compare_iteration(pod, {}, what=f"cornellSpaceship/sort {env}")


@pytest.mark.parametrize("tris,env", [(1000, {}), (20000, {}), (20000, {"B2PT_LONG_WALK": "2"})])
def test_two_meshes_one_scaled_bitexact(tmp_path, monkeypatch, tris, env):
    """Two OBJ geoms, the second scaled by 1.5 (not rigid: its object-space
    distances compete with world-space ones exactly as in the reference,
    SURVEY.md Q8): a lane walks the meshes of its ray one after the other;
    every stage identical to the oracle's loop over geoms and faces."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pod = _mesh_scene(tmp_path, "twoShips", 96, 54, tris)
    assert int((pod.geoms["type"] == abi.OBJ).sum()) == 2
    compare_iteration(pod, {}, what=f"twoShips/{tris}")
    d0 = oracle.intersect(pod, oracle.generate(pod, abi.default_options(), 1).origin,
                          oracle.generate(pod, abi.default_options(), 1).dir)
    hit_geoms = set(int(g) for g in d0.geom[d0.t > 0])
    obj_ids = [int(i) for i in np.nonzero(pod.geoms["type"] == abi.OBJ)[0]]
    assert all(g in hit_geoms for g in obj_ids), "both meshes must be visible"


def test_bvh_equals_brute_force_at_full_size(tmp_path):
    """Size-independent property at BASELINE.json's full size: at 1920x1080
    with the 250k-triangle stand-in mesh, the BVH kernel and the brute-force
    kernel (the reference's loop order) return identical hit records."""
    tris = int(os.environ.get("B2PT_TEST_TRIS", "250000"))
    root = assets.prepare(str(tmp_path / "run"), triangles=tris, procedural_size=512)
    pod = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, depth=2, root=root)).pod
    res = []
    for bvh in (1, 0):
        with api.Renderer(pod, abi.default_options(use_bvh=bvh, record_stages=1)) as r:
            r.render(1, 1, 1)
            st = r.stages()
            res.append((st, r.read()[0]))
    (a, img_a), (b, img_b) = res
    assert len(a) == len(b) == 2
    for d in range(2):
        for name in ["hit_t", "hit_normal", "hit_geom", "hit_face", "hit_uv", "hit_material", "sort_perm", "partition_pixel"]:
            assert_same_bits(a[d][name], b[d][name], f"depth {d} {name}")
    assert_same_bits(img_a, img_b, "image")
    n_obj = int((pod.geoms["type"][np.maximum(a[0]["hit_geom"], 0)][a[0]["hit_t"] > 0] == abi.OBJ).sum())
    assert n_obj > 10000


@pytest.mark.parametrize("scene,w,h,tris,iters", [("cornellSpaceship", 240, 135, 20000, 120), ("twoShips", 160, 90, 5000, 200)])
def test_bvh_never_loses_a_hit_soak(tmp_path, scene, w, h, tris, iters):
    """The BVH only prunes -- checked statistically: over >10 M path segments
    (many iterations, AA on, every depth) the accumulated image of the BVH walk
    is bit-identical to the brute-force kernel's (the reference's loop over every
    face).  One triangle culled wrongly by the padded FMA slab test, one tie
    resolved differently, would change a pixel."""
    pod = _mesh_scene(tmp_path, scene, w, h, tris)
    imgs, segs = [], []
    for bvh in (1, 0):
        with api.Renderer(pod, abi.default_options(use_bvh=bvh)) as r:
            r.render(1, iters, 1)
            imgs.append(r.read()[0])
            segs.append(int(r.live_counts()[:8].sum()))
    assert_same_bits(imgs[0], imgs[1], f"{scene}: BVH vs brute force after {iters} iterations")
    assert segs[0] == segs[1] and segs[0] * iters > 5_000_000, "too few segments for a soak test"


def test_multi_iteration_image_and_graph_replay(tmp_path):
    """50 iterations through the CUDA-graph path == 50 single launches == oracle."""
    pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=64, height=64)).pod
    ref, ref_alb, nlive, seg = oracle.render(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 50, 1)
    for graph in (1, 0):
        with api.Renderer(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE, use_graph=graph)) as r:
            r.render(1, 20, 1)
            r.render(21, 30, 1)
            img, alb = r.read()
            assert_same_bits(img, ref, f"image (graph={graph})")
            assert_same_bits(alb, ref_alb, "albedo")
            assert list(r.live_counts()[: len(nlive)]) == list(nlive)
            assert r.launch_count() > 50 * 3
            assert r.last_loop_ms() > 0.0


@pytest.mark.parametrize("scene_kind", ["cornellGlass", "mesh"])
def test_first_bounce_cache_changes_nothing(tmp_path, scene_kind):
    """a18 (pathtrace.cu:586-610, intended semantics): with AA and DOF off the
    depth-0 hits are computed once and reused; images, albedo, live counts and the
    per-depth stage dumps stay bit-identical to the uncached run and to the
    oracle, with and without the CUDA graph, and across a camera change."""
    if scene_kind == "mesh":
        pod = _mesh_scene(tmp_path, "cornellSpaceship", 96, 54, 1000)
    else:
        pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=64, height=48)).pod
    base = dict(trig_mode=abi.TRIG_PORTABLE, antialiasing=0)
    ref, ref_alb, nlive, _ = oracle.render(pod, abi.default_options(**base), 1, 6, 1)
    for graph in (1, 0):
        with api.Renderer(pod, abi.default_options(cache_first_bounce=1, use_graph=graph, **base)) as r:
            r.render(1, 1, 1)
            launches_fill = r.launch_count()
            r.render(2, 2, 1)
            r.render(4, 3, 1)
            img, alb = r.read()
            assert_same_bits(img, ref, f"cached image (graph={graph})")
            assert_same_bits(alb, ref_alb, "cached albedo")
            assert list(r.live_counts()[: len(nlive)]) == list(nlive)
            # iterations after the first skip the depth-0 intersect kernels
            per_iter_cached = (r.launch_count() - launches_fill - 2) / 5.0
            assert per_iter_cached < launches_fill - 1
    # stage dumps of a cached iteration == oracle stages of that iteration
    opt = abi.default_options(cache_first_bounce=1, record_stages=1, **base)
    oimg = np.zeros((pod.n_pixels, 3), np.float32)
    with api.Renderer(pod, opt) as r:
        for it in (1, 2, 3):
            r.render(it, 1, 1)
            got = r.stages()
            want = oracle.iteration_with_stages(pod, opt, it, oimg, None)
            for d, (g, o) in enumerate(zip(got, want)):
                for name in STAGES_TO_COMPARE:
                    assert_same_bits(g[name], o[name], f"cached iter {it} depth {d} {name}")
    # AA on: the option must be ignored (the reference compiles the cache out, pathtrace.cu:586)
    ref_aa, *_ = oracle.render(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 3, 1)
    with api.Renderer(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE, cache_first_bounce=1)) as r:
        r.render(1, 3, 1)
        assert_same_bits(r.read()[0], ref_aa, "cache ignored with AA on")


@pytest.mark.parametrize("env", [
    {"B2PT_LONG_WALK": "1"},                              # every walk finishes in the cooperative kernel (32 lanes per ray)
    {"B2PT_LONG_WALK": "1", "B2PT_LONG_LANES": "16"},     # ... in its 16-lane form (what contexts that share the SMs run)
    {"B2PT_LONG_WALK": "3", "B2PT_LONG_CARRY": "0"},      # ... restarting at the root, pruned by the carried hit
    {"B2PT_LONG_WALK": "3", "B2PT_LONG_CARRY": "0", "B2PT_LONG_LANES": "16"},
    {"B2PT_LONG_WALK": "2", "B2PT_LONG_CAP": "64"},       # hand-off queue full: lanes keep walking
    {"B2PT_LONG_WALK": "1000000"},                        # no hand-off at all
])
def test_long_walk_handoff_paths_bitexact(tmp_path, monkeypatch, env):
    """The hand-off of long BVH walks (k_mesh_walk -> k_mesh_walk_long) must not
    change a single bit whatever the threshold, carried-stack size or queue
    capacity: closest-hit ids, uv, normals and everything downstream against the
    oracle's brute-force loop."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 96, 54, 20000)
    compare_iteration(pod, {}, what=f"long walk {env}")
    with api.Renderer(pod, abi.default_options()) as r:
        r.render(1, 1, 1)
        walks, longs = int(r.walk_counts().sum()), int(r.walk_counts(True).sum())
    assert walks > 500
    if env["B2PT_LONG_WALK"] == "1":
        assert longs > walks // 4
    if env["B2PT_LONG_WALK"] == "1000000":
        assert longs == 0


@pytest.mark.parametrize("scene_kind", ["cornellGlass", "mesh", "twoShips"])
def test_fused_and_unfused_kernels_agree(tmp_path, monkeypatch, scene_kind):
    """k_generate_trace / k_shade_trace (analytic intersection fused into the kernel
    that produces the rays) against the separate kernels and the oracle: images,
    albedo and live counts bit-identical over several iterations, odd sizes included."""
    if scene_kind == "cornellGlass":
        pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=333, height=127)).pod
    else:
        pod = _mesh_scene(tmp_path, "cornellSpaceship" if scene_kind == "mesh" else "twoShips", 173, 99, 20000)
    ref, ref_alb, nlive, _ = oracle.render(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 5, 1)
    for fuse in ("1", "0"):
        monkeypatch.setenv("B2PT_FUSE", fuse)
        for graph in (1, 0):
            with api.Renderer(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE, use_graph=graph)) as r:
                r.render(1, 2, 1)
                r.render(3, 3, 1)
                img, alb = r.read()
                assert_same_bits(img, ref, f"{scene_kind} fuse={fuse} graph={graph} image")
                assert_same_bits(alb, ref_alb, f"{scene_kind} fuse={fuse} albedo")
                assert list(r.live_counts()[: len(nlive)]) == list(nlive)
                launches = r.launch_count()
        if fuse == "1":
            fused_launches = launches
    assert fused_launches < launches, "the fused form has fewer launches"


def test_contexts_are_independent_across_host_threads(tmp_path):
    """Two host threads, each creating, driving and destroying its own context
    (BVH build included) at the same time: both images equal the single-threaded one."""
    import threading

    pod = _mesh_scene(tmp_path, "cornellSpaceship", 128, 72, 5000)
    opt = abi.default_options(trig_mode=abi.TRIG_PORTABLE)
    with api.Renderer(pod, opt) as r:
        r.render(1, 12, 1)
        want, _ = r.read()
    out, errs = [None, None], []

    def work(k):
        try:
            for _ in range(3):  # create / destroy repeatedly while the other thread renders
                with api.Renderer(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE)) as rr:
                    rr.render(1, 12, 1)
                    out[k] = rr.read()[0]
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for k in range(2):
        assert_same_bits(out[k], want, f"thread {k}")


def test_shared_gpu_grids_change_nothing(tmp_path):
    """concurrent_contexts > 1 only resizes the persistent grids: every stage and
    the image stay bit-identical, also for four contexts rendering at once."""
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 160, 90, 20000)
    ref, _, nlive, _ = oracle.render(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 8, 1)
    compare_iteration(pod, {"concurrent_contexts": 4}, what="shared grids")
    opt = abi.default_options(trig_mode=abi.TRIG_PORTABLE, concurrent_contexts=4)
    rs = [api.Renderer(pod, opt) for _ in range(4)]
    try:
        for k, r in enumerate(rs):
            r.render(k + 1, 2, 4)           # iterations k+1 and k+5, all four contexts in flight
        total = np.zeros_like(ref)
        parts = [r.read()[0] for r in rs]
    finally:
        for r in rs:
            r.close()
    # each context's two iterations are bit-exact; their sum differs from the sequential sum only by float order
    for k, part in enumerate(parts):
        one, *_ = oracle.render(pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE), k + 1, 2, 4)
        assert_same_bits(part, one, f"context {k}")
        total += part
    assert np.allclose(total, ref, rtol=1e-5, atol=1e-6)


def test_strided_iterations_sum_to_the_sequential_image(tmp_path):
    """spp sharding: ranks rendering {1,3,5,..} and {2,4,6,..} add up to the
    sequential image within float summation order (1e-5 relative)."""
    pod = api.Scene(scenes.write_scene("cornell", str(tmp_path / "s.txt"), width=80, height=60)).pod
    opt = abi.default_options()
    with api.Renderer(pod, opt) as r:
        r.render(1, 16, 1)
        full, _ = r.read()
    parts = []
    for rank in range(2):
        with api.Renderer(pod, opt) as r:
            r.render(rank + 1, 8, 2)
            parts.append(r.read()[0])
    s = parts[0] + parts[1]
    assert np.allclose(s, full, rtol=1e-5, atol=1e-6)


def test_native_trig_converges_to_the_oracle_image(tmp_path):
    """NATIVE trig (CUDA sinf/cosf, as the reference's kernels) vs the libm
    oracle at equal spp: PSNR >= 50 dB on image/iter (north_star gate)."""
    pod = api.Scene(scenes.write_scene("cornell", str(tmp_path / "s.txt"), width=48, height=48)).pod
    n = 400
    with api.Renderer(pod, abi.default_options()) as r:
        r.render(1, n, 1)
        img, _ = r.read()
    ref, *_ = oracle.render(pod, abi.default_options(), 1, n, 1)
    assert psnr(np.clip(img / n, 0, 1), np.clip(ref / n, 0, 1)) >= 50.0


def test_pathtrace_free_functions_mirror_the_reference(tmp_path):
    scene = api.Scene(scenes.write_scene("cornell", str(tmp_path / "s.txt"), width=40, height=30))
    api.pathtraceFree()  # before the first Init, like main.cpp:245-248
    api.set_options(abi.default_options(trig_mode=abi.TRIG_PORTABLE))
    api.pathtraceInit(scene)
    for it in (1, 2, 3):
        api.pathtrace(None, 0, it)
        assert api.timer().getGpuElapsedTimeForPreviousOperation() > 0
    api.pathtraceFree()
    api.set_options(None)
    ref, alb, *_ = oracle.render(scene.pod, abi.default_options(trig_mode=abi.TRIG_PORTABLE), 1, 3, 1)
    assert_same_bits(scene.state.image, ref, "scene.state.image")
    assert_same_bits(scene.state.albedo, alb, "scene.state.albedo")


def test_edge_cases(tmp_path):
    # depth 1: every path dies at its first shade; depth 0 behaves like the reference's while loop
    for depth in (1, 2):
        pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / f"d{depth}.txt"), width=24, height=24, depth=depth)).pod
        compare_iteration(pod, {}, iters=(1,), what=f"depth {depth}")
    # a scene where every ray misses after the first bounce at most: the lone emissive sphere
    pod, *_ = load_golden("sphere_16x16")
    compare_iteration(pod, {}, iters=(1, 2, 3), what="sphere")
    # 1x1 image
    pod = api.Scene(scenes.write_scene("cornell", str(tmp_path / "one.txt"), width=1, height=1)).pod
    compare_iteration(pod, {}, iters=(1, 2), what="1x1")


@pytest.mark.parametrize("n_tris", [1, 2, 3, 4, 5, 8, 9, 17])
def test_tiny_meshes_bitexact(tmp_path, n_tris):
    """Meshes of a handful of triangles: up to kLeafTris the whole mesh is ONE leaf (no BVH node at all, the root
    is a leaf code), just above it the tree is a single wide node with leaves of uneven size."""
    (tmp_path / "models" / "materials").mkdir(parents=True)
    (tmp_path / "models" / "materials" / "tiny.mtl").write_text("newmtl plain\nKd 0.7 0.6 0.3\nKs 0.4 0.4 0.4\nKe 0 0 0\nNi 1.5\n")
    rng = np.random.default_rng(100 + n_tris)
    lines = ["mtllib tiny.mtl"]
    faces = []
    for k in range(n_tris):  # a fan of triangles facing the camera (+z seen from the ship's place in the box)
        c = rng.uniform(-1.2, 1.2, 3) * (1.0, 0.8, 0.3)
        a, b = rng.uniform(0.3, 0.9, 2)
        ang = rng.uniform(0, 2 * np.pi)
        p = [c + (a * np.cos(ang + t), b * np.sin(ang + t), 0.05 * k) for t in (0.0, 2.1, 4.2)]
        for q in p:
            lines.append("v %.6f %.6f %.6f" % tuple(q))
        faces.append((3 * k + 1, 3 * k + 2, 3 * k + 3))
    lines += ["vt 0.2 0.2", "vt 0.8 0.2", "vt 0.5 0.8", "vn 0 0 1", "usemtl plain"]
    back = n_tris not in (1, 3)  # 1 and 3 stay single sided: meshes of exactly 1 and 3 triangles
    for a, b, c in faces:
        lines.append(f"f {a}/1/1 {b}/2/1 {c}/3/1")
        if back:
            lines.append(f"f {a}/1/1 {c}/3/1 {b}/2/1")  # the back side, so that every ray direction finds a front face
    (tmp_path / "models" / "tiny.obj").write_text("\n".join(lines) + "\n")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=48, height=40, obj_path="../models/tiny.obj")
    pod = api.Scene(path).pod
    assert len(pod.face_pos) == (2 if back else 1) * n_tris
    compare_iteration(pod, {}, iters=(1, 2), what=f"{len(pod.face_pos)}-triangle mesh")


def test_many_geoms_and_materials(tmp_path):
    """Capacity corners: 64 geoms (the shared-memory staging and the 64-bit
    candidate mask of k_intersect_analytic) and 200 materials of all four BSDF
    kinds (200 populated bins in the one-pass material sort, look-back per bin)."""
    rng = np.random.default_rng(7)
    mats = [scenes._LIGHT]
    for i in range(199):
        kind = i % 4
        rgb = tuple(float(x) for x in rng.uniform(0.2, 0.95, 3).round(3))
        if kind == 0:
            mats.append((rgb, 0, (0, 0, 0), 0, 0, 0, 0))                      # diffuse
        elif kind == 1:
            mats.append((rgb, 0, rgb, 1, 0, 0, 0))                            # mirror
        elif kind == 2:
            mats.append((rgb, 0, rgb, 0, 1, round(float(rng.uniform(1.2, 1.8)), 2), 0))  # glass
        else:
            mats.append((rgb, 0, (0, 0, 0), 0, 0, 0, round(float(rng.uniform(0.5, 3.0)), 2)))  # small emitters
    objs = list(scenes._BOX)
    for i in range(58):
        kind = "sphere" if i % 2 else "cube"
        pos = tuple(float(x) for x in (rng.uniform(-4, 4), rng.uniform(0.5, 9), rng.uniform(-4, 4)))
        rot = tuple(float(x) for x in rng.uniform(0, 90, 3).round(1))
        scl = tuple(float(x) for x in rng.uniform(0.4, 1.3, 3).round(2))
        objs.append((kind, int(rng.integers(1, 200)), tuple(round(v, 2) for v in pos), rot, scl))
    assert len(objs) == 64
    scenes.SCENES["_capacity"] = dict(file="cornell", materials=mats, objects=objs)
    try:
        pod = api.Scene(scenes.write_scene("_capacity", str(tmp_path / "cap.txt"), width=96, height=64)).pod
    finally:
        del scenes.SCENES["_capacity"]
    assert len(pod.geoms) == 64 and len(pod.materials) == 200
    compare_iteration(pod, {}, iters=(1, 2), what="64 geoms / 200 materials")
    compare_iteration(pod, {"sort_by_material": 0}, iters=(1,), what="64 geoms, no sort")
    # one geom too many is refused, not truncated
    scenes.SCENES["_over"] = dict(file="cornell", materials=mats, objects=objs + [objs[-1]])
    try:
        over = api.Scene(scenes.write_scene("_over", str(tmp_path / "over.txt"), width=8, height=8)).pod
    finally:
        del scenes.SCENES["_over"]
    with pytest.raises(api.B2ptError):
        api.Renderer(over)


def test_invalid_scenes_are_rejected():
    pod, *_ = load_golden("cornell_32x32")
    bad = pod.copy()
    bad.geoms["material_id"][0] = 99
    with pytest.raises(api.B2ptError):
        api.Renderer(bad)
    bad = pod.copy()
    bad.trace_depth = 1000
    with pytest.raises(api.B2ptError):
        api.Renderer(bad)


# ---- the pixel-keyed RNG (north_star's second mode: a counter keyed on (pixel, iteration, depth)) ---------------
@pytest.mark.parametrize("case,optkw", [("cornellGlass_32x32", {}), ("both_refl_refr_48x20", {}), ("texquad_32x32", {}),
                                         ("cornell_32x32", {"sort_by_material": 0})])
def test_pixel_rng_every_stage_bitexact(case, optkw):
    """rng_mode = PIXEL against the oracle's restatement of it, every stage of two iterations."""
    pod, *_ = load_golden(case)
    compare_iteration(pod, dict(rng_mode=abi.RNG_PIXEL, **optkw), iters=(1, 2), what=f"{case} pixel rng")


def test_pixel_rng_mesh_scene_and_order_independence(tmp_path):
    """With pixel keys the radiance of a path no longer depends on its slot: the image is the same, bit for bit,
    with and without the material sort, from one context or from members that share the iterations."""
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 160, 90, 5000)
    compare_iteration(pod, dict(rng_mode=abi.RNG_PIXEL), iters=(1,), what="mesh scene, pixel rng")
    imgs = []
    for sort in (1, 0):
        with api.Renderer(pod, abi.default_options(rng_mode=abi.RNG_PIXEL, sort_by_material=sort)) as r:
            r.render(1, 6, 1)
            imgs.append(r.read()[0])
    assert_same_bits(imgs[0], imgs[1], "pixel rng: material sort on / off")
    n = pod.n_pixels
    host = np.zeros((n, 3), np.float32)
    with api.MultiRenderer(pod, abi.default_options(rng_mode=abi.RNG_PIXEL, sort_by_material=0), devices=[0, 0, 0], lanes=2) as multi:
        multi.pathtrace(1, host, None)
        multi.pathtrace(4, host, None)
    assert_same_bits(imgs[0], host, "pixel rng: three members, two frames")
    # and it is a different random sequence from the reference's slot keys
    with api.Renderer(pod, abi.default_options()) as r:
        r.render(1, 6, 1)
        slot = r.read()[0]
    assert not np.array_equal(slot, imgs[0])


def test_pixel_rng_converges_to_the_slot_keyed_image(tmp_path):
    """north_star's gate for the second RNG mode: the converged image at equal spp reaches PSNR >= 50 dB against the
    reference-mode (slot-keyed) image.  Two unbiased estimates with independent noise differ by MSE = 2 var / N, so the
    PSNR must ALSO grow by ~3 dB per doubling of N -- a biased mode would level off instead.  The scene is the Cornell
    box with the whole ceiling as a dim light (per-sample variance ~0.03, so that 50 dB is reached at ~10^4 spp; the
    shipped scenes' small bright light needs ~10^6 spp for the same PSNR, with any RNG)."""
    mats = [((1, 1, 1), 0, (0, 0, 0), 0, 0, 0, 0.25), scenes._WHITE, scenes._RED, scenes._GREEN, scenes._MIRROR, scenes._GLASS]
    objs = [("cube", 0, (0, 10, 0), (0, 0, 0), (10, .3, 10))] + list(scenes._BOX[1:]) + [("sphere", 5, (-1, 4, -1), (0, 0, 0), (3, 3, 3))]
    scenes.SCENES["_bright"] = dict(file="cornell", materials=mats, objects=objs)
    try:
        pod = api.Scene(scenes.write_scene("_bright", str(tmp_path / "s.txt"), width=40, height=40)).pod
    finally:
        del scenes.SCENES["_bright"]

    def render(mode, sort, spp):
        with api.Renderer(pod, abi.default_options(rng_mode=mode, sort_by_material=sort)) as r:
            r.render(1, spp, 1)
            return np.clip(r.read()[0] / np.float32(spp), 0.0, 1.0)

    got = {}
    for spp in (2048, 32768):
        got[spp] = psnr(render(abi.RNG_SLOT, 1, spp), render(abi.RNG_PIXEL, 0, spp))
    assert got[32768] >= 50.0, got
    assert 9.0 <= got[32768] - got[2048] <= 15.0, got  # 16x the samples: +12 dB


# ---- per-face materials (B2ptScene::face_material / material_textures; the reference discards the ids) -----------
def test_per_face_materials_small_obj_bitexact(tmp_path):
    """tests/golden/multimat.obj through the loader with per_face_materials: faces of one OBJ shaded with different
    materials, one of them with all four maps; every stage against the oracle, BVH walk and brute force."""
    import shutil

    (tmp_path / "models" / "materials").mkdir(parents=True)
    (tmp_path / "textures").mkdir()
    shutil.copy(os.path.join(GOLDEN, "multimat.obj"), tmp_path / "models")
    shutil.copy(os.path.join(GOLDEN, "multimat.mtl"), tmp_path / "models" / "materials")
    for f in os.listdir(os.path.join(GOLDEN, "texquad")):
        shutil.copy(os.path.join(GOLDEN, "texquad", f), tmp_path / "textures")
    path = scenes.write_scene("cornellObj", str(tmp_path / "scenes" / "s.txt"), width=64, height=48, obj_path="../models/multimat.obj")
    pod = api.Scene(path, per_face_materials=True).pod
    assert pod.face_material is not None and len(set(pod.face_material.tolist())) >= 3
    compare_iteration(pod, {}, iters=(1, 2), what="multimat")
    compare_iteration(pod, {"use_bvh": 0}, iters=(1,), what="multimat, brute force")
    compare_iteration(pod, {"sort_by_material": 0}, iters=(1,), what="multimat, no sort")


@pytest.mark.parametrize("env", [{}, {"B2PT_LONG_WALK": "2"}])
def test_per_face_materials_large_mesh_bitexact(tmp_path, monkeypatch, env):
    """The stand-in hull with its faces spread over four materials (one keeps the maps, one is emissive) in two OBJ
    geoms: the fold of the walk, the long-walk kernel and the material histograms all work per face."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pod = _mesh_scene(tmp_path, "twoShips", 112, 63, 5000).copy()
    base = len(pod.materials)
    extra = np.zeros(3, pod.materials.dtype)
    extra["color"] = [(0.2, 0.7, 0.9), (0.9, 0.9, 0.2), (1.0, 0.8, 0.6)]
    extra["specular_color"] = [(0.5, 0.5, 0.5), (0.9, 0.2, 0.2), (0, 0, 0)]
    extra["index_of_refraction"] = [1.3, 2.2, 1.0]
    extra["emittance"] = [0.0, 0.0, 2.0]
    pod.materials = np.concatenate([pod.materials, extra])
    rng = np.random.default_rng(5)
    fm = np.zeros(len(pod.face_pos), np.int32)
    mt = np.full((len(pod.materials), 4), -1, np.int32)
    for g in np.nonzero(pod.geoms["type"] == abi.OBJ)[0]:
        fb, fc, gm = int(pod.geoms["face_begin"][g]), int(pod.geoms["face_count"][g]), int(pod.geoms["material_id"][g])
        choice = rng.integers(0, 4, fc)
        fm[fb: fb + fc] = np.where(choice == 0, gm, base + choice - 1)
        mt[gm] = [pod.geoms[k][g] for k in ("tex_kd", "tex_ks", "tex_bump", "tex_ke")]
    mt[base] = mt[int(pod.geoms["material_id"][np.nonzero(pod.geoms["type"] == abi.OBJ)[0][0]])]  # a second material with the maps
    pod.face_material, pod.material_textures = fm, mt
    compare_iteration(pod, {}, iters=(1, 2), what=f"per-face materials {env}")

"""Helpers shared by the tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def bits(a):
    a = np.ascontiguousarray(np.squeeze(a))
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same_bits(a, b) -> bool:
    a, b = bits(a), bits(b)
    return a.shape == b.shape and bool(np.array_equal(a, b))


def assert_same_bits(a, b, what=""):
    a2, b2 = bits(a), bits(b)
    assert a2.shape == b2.shape, f"{what}: shape {a2.shape} vs {b2.shape}"
    bad = int((a2 != b2).sum())
    assert bad == 0, f"{what}: {bad} of {a2.size} elements differ bitwise"


def load_golden(case):
    from mygpuraytracer_b200.podscene import PodScene

    scene = PodScene.load(os.path.join(GOLDEN, case + ".b2s"))
    z = np.load(os.path.join(GOLDEN, case + "_stages.npz"))
    depths = []
    k = 0
    while f"d{k}_hit_t" in z:
        depths.append({n[len(f"d{k}_"):]: z[n] for n in z.files if n.startswith(f"d{k}_")})
        k += 1
    return scene, depths, z["image"], z["albedo"], z["nlive"]


# oracle stage name -> golden (reference dump) stage name
STAGE_MAP = [
    ("ray_origin", "in_origin"), ("ray_dir", "in_dir"), ("ray_pixel", "in_pixel"), ("hit_t", "hit_t"),
    ("hit_normal", "hit_normal"), ("hit_material", "hit_material"), ("sorted_pixel", "sorted_pixel"),
    ("shaded_color", "shaded_color"), ("shaded_bounces", "shaded_bounces"), ("shaded_origin", "shaded_origin"),
    ("shaded_dir", "shaded_dir"), ("partition_pixel", "part_pixel"),
]


def psnr(a, b, peak=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)

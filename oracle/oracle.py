"""ctypes wrapper of ``oracle/libpt_oracle.so``.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs import
this module.  It wraps the plain-C restatement of the reference algorithm
(``oracle/pt_oracle.c``) stage by stage so a test can drive the wavefront loop
and keep every intermediate array.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional

import numpy as np

from mygpuraytracer_b200 import abi
from mygpuraytracer_b200.podscene import PodScene

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpt_oracle.so")
_lib: Optional[C.CDLL] = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, seconds)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "pt_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.oracle_utilhash.restype = C.c_uint32
        L.oracle_utilhash.argtypes = [C.c_uint32]
        L.oracle_seed.restype = C.c_uint32
        L.oracle_seed.argtypes = [i32, i32, i32]
        L.oracle_minstd_next.restype = C.c_uint32
        L.oracle_minstd_next.argtypes = [C.POINTER(C.c_uint32)]
        L.oracle_uniform.restype = f32
        L.oracle_uniform.argtypes = [C.POINTER(C.c_uint32), f32, f32]
        for n in ("oracle_box_test", "oracle_sphere_test"):
            getattr(L, n).restype = f32
            getattr(L, n).argtypes = [C.POINTER(abi.Geom), vp, vp, vp]
        L.oracle_ray_triangle.restype = C.c_int
        L.oracle_ray_triangle.argtypes = [vp] * 6
        L.oracle_hemisphere.restype = None
        L.oracle_hemisphere.argtypes = [vp, C.POINTER(C.c_uint32), i32, vp]
        L.oracle_sincos_portable.restype = None
        L.oracle_sincos_portable.argtypes = [f32, C.POINTER(f32), C.POINTER(f32)]
        L.oracle_generate.restype = C.c_int
        L.oracle_generate.argtypes = [C.POINTER(abi.Camera), C.POINTER(abi.Options), i32, i32] + [vp] * 5
        L.oracle_intersect.restype = C.c_int
        L.oracle_intersect.argtypes = [C.POINTER(abi.Scene), i32] + [vp] * 8
        L.oracle_sort_perm.restype = C.c_int
        L.oracle_sort_perm.argtypes = [i32, vp, vp]
        L.oracle_shade.restype = C.c_int
        L.oracle_shade.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Options), i32, i32, i32] + [vp] * 11
        L.oracle_partition_perm.restype = C.c_int
        L.oracle_partition_perm.argtypes = [i32, vp, vp]
        L.oracle_gather.restype = C.c_int
        L.oracle_gather.argtypes = [i32, vp, vp, vp]
        L.oracle_render.restype = C.c_int
        L.oracle_render.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Options), i32, i32, i32, vp, vp, vp, vp]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_set_dof_arg_order.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


# ---- scalar helpers -----------------------------------------------------------------
def seed(iter_: int, index: int, depth: int) -> int:
    return lib().oracle_seed(iter_, index, depth)


def draws(iter_: int, index: int, depth: int, n: int, a: float = 0.0, b: float = 1.0) -> List[float]:
    st = C.c_uint32(seed(iter_, index, depth))
    return [float(lib().oracle_uniform(C.byref(st), a, b)) for _ in range(n)]


def minstd_nth(n: int, seed_value: int = 1) -> int:
    st = C.c_uint32(seed_value)
    v = 0
    for _ in range(n):
        v = lib().oracle_minstd_next(C.byref(st))
    return v


def hemisphere(normal, iter_: int, index: int, depth: int, trig_mode: int = abi.TRIG_NATIVE) -> np.ndarray:
    st = C.c_uint32(seed(iter_, index, depth))
    n = np.asarray(normal, np.float32)
    out = np.zeros(3, np.float32)
    lib().oracle_hemisphere(_p(n), C.byref(st), trig_mode, _p(out))
    return out


def sincos_portable(x: np.ndarray):
    x = np.asarray(x, np.float32).ravel()
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    fs, fc = C.c_float(), C.c_float()
    for i, v in enumerate(x):
        lib().oracle_sincos_portable(float(v), C.byref(fs), C.byref(fc))
        s[i], c[i] = fs.value, fc.value
    return s, c


def geom_test(kind: str, geom: abi.Geom, origin, direction):
    o = np.asarray(origin, np.float32)
    d = np.asarray(direction, np.float32)
    n = np.zeros(3, np.float32)
    fn = lib().oracle_box_test if kind == "box" else lib().oracle_sphere_test
    t = fn(C.byref(geom), _p(o), _p(d), _p(n))
    return float(t), n


def ray_triangle(o, d, v0, v1, v2):
    arrs = [np.asarray(a, np.float32) for a in (o, d, v0, v1, v2)]
    bary = np.zeros(3, np.float32)
    hit = lib().oracle_ray_triangle(*[_p(a) for a in arrs], _p(bary))
    return bool(hit), bary


# ---- stages ----------------------------------------------------------------------------
class Paths:
    """SoA path state in slot order (PathSegment, apps/src/sceneStructs.h:105-110)."""

    def __init__(self, n: int):
        self.origin = np.zeros((n, 3), np.float32)
        self.dir = np.zeros((n, 3), np.float32)
        self.color = np.zeros((n, 3), np.float32)
        self.pixel = np.zeros(n, np.int32)
        self.bounces = np.zeros(n, np.int32)

    def __len__(self):
        return len(self.pixel)

    def take(self, idx: np.ndarray) -> "Paths":
        p = Paths(0)
        p.origin, p.dir, p.color = self.origin[idx].copy(), self.dir[idx].copy(), self.color[idx].copy()
        p.pixel, p.bounces = self.pixel[idx].copy(), self.bounces[idx].copy()
        return p

    def copy(self) -> "Paths":
        return self.take(np.arange(len(self)))


class Hits:
    """SoA hit records (ShadeableIntersection, apps/src/sceneStructs.h:115-121) + face id."""

    def __init__(self, n: int):
        self.t = np.zeros(n, np.float32)
        self.normal = np.zeros((n, 3), np.float32)
        self.uv = np.zeros((n, 2), np.float32)
        self.geom = np.zeros(n, np.int32)
        self.face = np.zeros(n, np.int32)
        self.material = np.zeros(n, np.int32)

    def take(self, idx: np.ndarray) -> "Hits":
        h = Hits(0)
        h.t, h.normal, h.uv = self.t[idx].copy(), self.normal[idx].copy(), self.uv[idx].copy()
        h.geom, h.face, h.material = self.geom[idx].copy(), self.face[idx].copy(), self.material[idx].copy()
        return h


def generate(scene: PodScene, opt: abi.Options, iter_: int) -> Paths:
    cs = scene.as_ctypes()
    p = Paths(scene.n_pixels)
    rc = lib().oracle_generate(C.byref(cs.camera), C.byref(opt), iter_, scene.trace_depth, _p(p.origin), _p(p.dir),
                               _p(p.color), _p(p.pixel), _p(p.bounces))
    assert rc == 0
    return p


def intersect(scene: PodScene, origin: np.ndarray, direction: np.ndarray) -> Hits:
    cs = scene.as_ctypes()
    o = np.ascontiguousarray(origin, np.float32)
    d = np.ascontiguousarray(direction, np.float32)
    n = len(o)
    h = Hits(n)
    rc = lib().oracle_intersect(C.byref(cs), n, _p(o), _p(d), _p(h.t), _p(h.normal), _p(h.uv), _p(h.geom), _p(h.face),
                                _p(h.material))
    assert rc == 0
    return h


def sort_perm(material: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(material, np.int32)
    perm = np.zeros(len(m), np.int32)
    rc = lib().oracle_sort_perm(len(m), _p(m), _p(perm))
    assert rc == 0
    return perm


def shade(scene: PodScene, opt: abi.Options, iter_: int, depth: int, hits: Hits, paths: Paths,
          albedo: Optional[np.ndarray]) -> None:
    """In-place shade of `paths` (both arguments in SORTED slot order)."""
    cs = scene.as_ctypes()
    n = len(paths)
    rc = lib().oracle_shade(C.byref(cs), C.byref(opt), iter_, depth, n, _p(hits.t), _p(hits.normal), _p(hits.uv),
                            _p(hits.geom), _p(hits.material), _p(paths.origin), _p(paths.dir), _p(paths.color),
                            _p(paths.pixel), _p(paths.bounces), _p(albedo) if albedo is not None else None)
    assert rc == 0


def partition_perm(bounces: np.ndarray):
    b = np.ascontiguousarray(bounces, np.int32)
    perm = np.zeros(len(b), np.int32)
    live = lib().oracle_partition_perm(len(b), _p(b), _p(perm))
    return perm, live


def gather(image: np.ndarray, color: np.ndarray, pixel: np.ndarray) -> None:
    c = np.ascontiguousarray(color, np.float32)
    p = np.ascontiguousarray(pixel, np.int32)
    lib().oracle_gather(len(p), _p(image), _p(c), _p(p))


def iteration_with_stages(scene: PodScene, opt: abi.Options, iter_: int, image: np.ndarray,
                          albedo: Optional[np.ndarray]) -> List[Dict[str, np.ndarray]]:
    """One iteration of apps/src/pathtrace.cu:572-655 driven stage by stage;
    returns one dict of stage arrays per depth (names as in abi.STAGES, plus
    ``ray_*`` for the rays entering the depth) and accumulates into image."""
    paths = generate(scene, opt, iter_)
    final = paths.copy()  # the full P-long array finalGather sees
    n = len(paths)
    live_idx = np.arange(n)  # where the live prefix sits inside `final` (always the prefix)
    stages = []
    depth = 0
    while n > 0:
        st: Dict[str, np.ndarray] = {}
        st["ray_origin"], st["ray_dir"], st["ray_pixel"] = paths.origin.copy(), paths.dir.copy(), paths.pixel.copy()
        st["ray_color"], st["ray_bounces"] = paths.color.copy(), paths.bounces.copy()
        hits = intersect(scene, paths.origin, paths.dir)
        st.update(hit_t=hits.t, hit_normal=hits.normal, hit_uv=hits.uv, hit_geom=hits.geom, hit_face=hits.face,
                  hit_material=hits.material)
        if opt.sort_by_material:
            perm = sort_perm(hits.material)
        else:
            perm = np.arange(n, dtype=np.int32)
        st["sort_perm"] = perm
        hits_s, paths = hits.take(perm), paths.take(perm)
        st["sorted_pixel"] = paths.pixel.copy()
        depth += 1
        shade(scene, opt, iter_, depth, hits_s, paths, albedo)
        st["shaded_color"], st["shaded_bounces"] = paths.color.copy(), paths.bounces.copy()
        st["shaded_origin"], st["shaded_dir"] = paths.origin.copy(), paths.dir.copy()
        pperm, live = partition_perm(paths.bounces)
        paths = paths.take(pperm)
        st["partition_pixel"] = paths.pixel.copy()
        st["n_live_out"] = np.int32(live)
        # write the partitioned block back over the prefix of the full array
        for name in ("origin", "dir", "color", "pixel", "bounces"):
            getattr(final, name)[:n] = getattr(paths, name)
        paths = paths.take(np.arange(live))
        n = live
        stages.append(st)
    del live_idx
    gather(image, final.color, final.pixel)
    return stages


def render(scene: PodScene, opt: abi.Options, iter_first: int = 1, count: int = 1, stride: int = 1,
           image: Optional[np.ndarray] = None, albedo: Optional[np.ndarray] = None):
    """Whole iterations in C (OpenMP).  Returns (image, albedo, n_live, segments)."""
    cs = scene.as_ctypes()
    P = scene.n_pixels
    if image is None:
        image = np.zeros((P, 3), np.float32)
    if albedo is None:
        albedo = np.zeros((P, 3), np.float32)
    n_live = np.zeros(scene.trace_depth + 1, np.int32)
    seg = C.c_int64(0)
    rc = lib().oracle_render(C.byref(cs), C.byref(opt), iter_first, count, stride, _p(image), _p(albedo), _p(n_live),
                             C.addressof(seg))
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    return image, albedo, n_live, int(seg.value)


def set_dof_arg_order(right_to_left: bool) -> None:
    lib().oracle_set_dof_arg_order(1 if right_to_left else 0)


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(n)


# ---- output hand-off (TEST INFRASTRUCTURE like everything in this package) -----------
def save_image_rgb8(image: np.ndarray, width: int, height: int, samples: float = 1.0, divide: bool = True,
                    mirror_x: bool = True) -> np.ndarray:
    """saveImage + image::savePNG quantisation (apps/src/main.cpp:115-135,
    apps/src/image.cpp:22-33): ``img.setPixel(width-1-x, y, pix / samples)``,
    then ``(unsigned char)(glm::clamp(pix, 0, 1) * 255.f)`` per channel.
    Returns (H, W, 3) uint8."""
    v = np.asarray(image, np.float32).reshape(height, width, 3)
    if divide:
        v = (v / np.float32(samples)).astype(np.float32)
    c = np.where(v < np.float32(0), np.float32(0), v)          # glm::max(x, 0) = (x < 0) ? 0 : x
    c = np.where(np.float32(1) < c, np.float32(1), c)          # glm::min(x, 1) = (1 < x) ? 1 : x
    c = (c * np.float32(255.0)).astype(np.float32)
    c = np.where(np.isnan(c), np.float32(0), c)                # the reference's cast of NaN is undefined
    out = c.astype(np.uint8)                                   # truncation, values already in [0, 255]
    return out[:, ::-1, :].copy() if mirror_x else out


def save_image_rgbe(image: np.ndarray, width: int, height: int, samples: float = 1.0, divide: bool = True,
                    mirror_x: bool = True) -> np.ndarray:
    """saveImage + image::saveHDR (apps/src/main.cpp:115-135,163, apps/src/image.cpp:41-45): the pixels are
    mirrored and divided as for the PNG, then stb_image_write's Radiance encoder turns each into shared-exponent
    RGBE (Ward's float2rgbe as stbi_write_hdr states it): with m the largest channel, a pixel below 1e-32 is
    four zero bytes, otherwise ``frexp(m) = (f, e)``, ``scale = float(f) * 256.0f / m`` in single precision,
    mantissas ``(unsigned char)(channel * scale)`` (truncation) and exponent byte ``e + 128``.
    Returns (H, W, 4) uint8.  Pinned against the reference's own image.cpp (tests/golden/hdr_golden.npz)."""
    v = np.asarray(image, np.float32).reshape(height, width, 3)
    if divide:
        v = (v / np.float32(samples)).astype(np.float32)
    if mirror_x:
        v = v[:, ::-1, :]
    m = np.maximum(v[..., 0], np.maximum(v[..., 1], v[..., 2])).astype(np.float32)
    live = ~(m.astype(np.float64) < 1e-32)
    f, e = np.frexp(np.where(live, m, np.float32(1)))
    scale = (f.astype(np.float32) * np.float32(256.0)).astype(np.float32)
    scale = (scale / np.where(live, m, np.float32(1))).astype(np.float32)
    out = np.zeros((height, width, 4), np.uint8)
    for c in range(3):
        out[..., c] = (v[..., c] * scale).astype(np.float32).astype(np.int32).astype(np.uint8)
    out[..., 3] = (e + 128).astype(np.uint8)
    out[~live] = 0
    return out


def denoise_color(image: np.ndarray, iteration: int) -> np.ndarray:
    """inputColor[index] = image[index] / (float)iteration (CPUdenoise, apps/src/main.cpp:194-199)."""
    return (np.asarray(image, np.float32) / np.float32(iteration)).astype(np.float32)

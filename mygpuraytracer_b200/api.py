"""Host-side mirror of the reference's render API over the C ABI.

The reference exposes five free functions with file-static state
(apps/src/pathtrace.h:6-10) and a ``Scene`` class with public vectors
(apps/src/scene.h:12-32).  This module keeps those names and their meaning:

========================  =====================================================
reference                 here
========================  =====================================================
``new Scene(path)``       :class:`Scene` (``scene.geoms``, ``.materials``,
                          ``.state.image``, ``.state.albedo``, ...)
``pathtraceInit(scene)``  :func:`pathtraceInit`
``pathtrace(pbo, f, it)`` :func:`pathtrace` -- renders iteration ``it`` and
                          copies the running sum and the albedo AOV into
                          ``scene.state.image`` / ``.albedo`` exactly as
                          apps/src/pathtrace.cu:663-668 does
``pathtraceFree()``       :func:`pathtraceFree` (safe before the first Init,
                          apps/src/main.cpp:245-248)
``timer()``               :func:`timer` -> object with
                          ``getGpuElapsedTimeForPreviousOperation()``
``sendToGPU(pbo, iter)``  :func:`sendToGPU` (tone-map into an RGBA8 buffer)
========================  =====================================================

:class:`Renderer` is the explicit-context form of the same calls (several
contexts, one per GPU, can coexist; it also exposes the batched
``render(iter_first, count, stride)`` used for samples-per-pixel sharding).

Everything computes on the GPU through ``libb2pt.so``.  There is no CPU
fallback: if the library has not been built, or no CUDA device is present,
the calls raise :class:`B2ptError`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import numpy as np

from . import abi
from .podscene import PodScene

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2PT_LIB: an alternative build of the same library (kernel experiments, tools/build_variants.py)
LIB_PATH = os.environ.get("B2PT_LIB") or os.path.join(_HERE, "libb2pt.so")
_lib: Optional[C.CDLL] = None


class B2ptError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b2pt error {code} ({abi.ERR_NAMES.get(code, '?')}): {message}")
        self.code = code


def load_library() -> C.CDLL:
    """dlopen ``libb2pt.so`` and check every symbol of ``include/b2pt.h``."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2ptError(abi.B2PT_OK - 5, f"{LIB_PATH} is missing: run `python -m mygpuraytracer_b200.build` "
                                             "(there is no CPU fallback)")
        _lib = abi.declare(C.CDLL(LIB_PATH))
    return _lib


def _check(rc: int) -> int:
    if rc < 0:
        raise B2ptError(rc, load_library().b2pt_last_error().decode("utf-8", "replace"))
    return rc


def device_count() -> int:
    return _check(load_library().b2pt_device_count())


# ---------------------------------------------------------------------------------------
# Scene
# ---------------------------------------------------------------------------------------
class RenderState:
    """RenderState, apps/src/sceneStructs.h:95-103."""

    def __init__(self, pod: PodScene, image_name: str):
        self.camera = pod.camera
        self.iterations = pod.iterations
        self.traceDepth = pod.trace_depth
        self.imageName = image_name
        n = pod.n_pixels
        self.image = np.zeros((n, 3), np.float32)
        self.albedo = np.zeros((n, 3), np.float32)
        self.output = np.zeros((n, 3), np.float32)


class Scene:
    """``Scene(filename)``, apps/src/scene.cpp:10-36, via ``b2pt_scene_load``.

    ``width``/``height``/``iterations``/``depth`` override the RES /
    ITERATIONS / DEPTH lines (every shipped scene says 800x800, 5000, 8).
    ``per_face_materials`` keeps tinyobj's per-face material ids, which the
    reference reads and discards (apps/src/scene.cpp:121-122).
    A :class:`Scene` can also wrap an existing :class:`PodScene`.
    """

    def __init__(self, filename: Optional[str] = None, *, width: int = 0, height: int = 0, iterations: int = 0,
                 depth: int = 0, per_face_materials: bool = False, pod: Optional[PodScene] = None):
        if pod is None:
            if filename is None:
                raise ValueError("Scene needs a file name or a PodScene")
            lib = load_library()
            ov = abi.LoadOverrides(width, height, iterations, depth, 1 if per_face_materials else 0)
            h = C.c_void_p()
            _check(lib.b2pt_scene_load(os.fsencode(filename), C.byref(ov), C.byref(h)))
            try:
                pod = PodScene.from_ctypes(lib.b2pt_scene_view(h).contents)
                name = lib.b2pt_scene_image_name(h).decode("utf-8", "replace")
                warnings = [w for w in lib.b2pt_scene_warnings(h).decode("utf-8", "replace").splitlines() if w]
            finally:
                lib.b2pt_scene_free(h)
        else:
            name, warnings = "pod", []
        self.warnings = warnings  # what the loader tolerated like the reference does (missing / undecodable textures)
        self.pod = pod
        self.state = RenderState(pod, name)

    # the reference's public vectors
    @property
    def geoms(self):
        return self.pod.geoms

    @property
    def materials(self):
        return self.pod.materials

    @property
    def allFaces(self):
        return [self.pod.face_pos[g["face_begin"]: g["face_begin"] + g["face_count"]] for g in self.pod.geoms]


# ---------------------------------------------------------------------------------------
# Renderer: one context
# ---------------------------------------------------------------------------------------
class Renderer:
    def __init__(self, scene, options: Optional[abi.Options] = None, share: Optional["Renderer"] = None, **opt_kw):
        """``share``: a Renderer of the SAME scene on the same GPU whose device copy of the scene (BVH,
        triangles, textures) this context uses instead of uploading its own (``b2pt_create_shared``)."""
        self.lib = load_library()
        self.pod: PodScene = scene.pod if isinstance(scene, Scene) else scene
        self.options = options if options is not None else abi.default_options(**opt_kw)
        self._cscene = self.pod.as_ctypes()
        self._h = C.c_void_p()
        if share is not None:
            _check(self.lib.b2pt_create_shared(share._h, C.byref(self._cscene), C.byref(self.options), C.byref(self._h)))
        else:
            _check(self.lib.b2pt_create(C.byref(self._cscene), C.byref(self.options), C.byref(self._h)))
        self.n_pixels = self.pod.n_pixels

    @classmethod
    def view(cls, handle, pod: PodScene) -> "Renderer":
        """A non-owning Renderer over an existing ``B2ptCtx*`` (a lane of a Pipeline / Shard): statistics and
        per-kernel profiling of the contexts that really render.  ``close()`` does not destroy the context."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.pod = pod
        self.options = None
        self._cscene = None
        self._h = C.c_void_p(handle if isinstance(handle, int) else C.cast(handle, C.c_void_p).value)
        self._borrowed = True
        self.n_pixels = pod.n_pixels
        return self

    # -- lifetime ------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            if not getattr(self, "_borrowed", False):
                self.lib.b2pt_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    # -- rendering -----------------------------------------------------------------------
    def render(self, iter_first: int = 1, count: int = 1, stride: int = 1) -> None:
        _check(self.lib.b2pt_render(self._h, iter_first, count, stride))

    def sync(self) -> None:
        _check(self.lib.b2pt_sync(self._h))

    def read(self, image: Optional[np.ndarray] = None, albedo: Optional[np.ndarray] = None, want_albedo: bool = True):
        if image is None:
            image = np.empty((self.n_pixels, 3), np.float32)
        if albedo is None and want_albedo:
            albedo = np.empty((self.n_pixels, 3), np.float32)
        _check(self.lib.b2pt_read_accum(self._h, image.ctypes.data, albedo.ctypes.data if albedo is not None else None))
        return image, albedo

    def pathtrace(self, iteration: int, image: np.ndarray, albedo: Optional[np.ndarray]) -> None:
        """One reference ``pathtrace()`` call.  ``albedo`` is written on every call unless the context was created
        with ``persistent_host_albedo=1``: then the caller vouches that it passes the same, otherwise untouched
        array every time, and an unchanged AOV is not copied again."""
        _check(self.lib.b2pt_pathtrace(self._h, iteration, image.ctypes.data,
                                       albedo.ctypes.data if albedo is not None else None))

    def reset(self) -> None:
        _check(self.lib.b2pt_reset_accum(self._h))

    def set_camera(self, camera: np.ndarray) -> None:
        cam = abi.Camera()
        C.memmove(C.byref(cam), np.ascontiguousarray(camera).ctypes.data, C.sizeof(abi.Camera))
        _check(self.lib.b2pt_set_camera(self._h, C.byref(cam)))

    # -- output hand-off (saveImage / CPUdenoise, apps/src/main.cpp:115-219) ---------------
    def resolve_rgb8(self, aov: int = abi.AOV_IMAGE, samples: int = 1, mirror_x: bool = True) -> np.ndarray:
        """The (H, W, 3) uint8 pixels ``saveImage`` hands to ``image::savePNG``."""
        out = np.empty((self.pod.height, self.pod.width, 3), np.uint8)
        _check(self.lib.b2pt_resolve_rgb8(self._h, aov, samples, 1 if mirror_x else 0, out.ctypes.data))
        return out

    def save_png(self, path: str, aov: int = abi.AOV_IMAGE, samples: int = 1) -> None:
        _check(self.lib.b2pt_save_png(self._h, aov, samples, os.fsencode(path)))

    def save_hdr(self, path: str, aov: int = abi.AOV_IMAGE, samples: int = 1) -> None:
        """``saveImage`` + ``image::saveHDR`` (apps/src/image.cpp:41-45)."""
        _check(self.lib.b2pt_save_hdr(self._h, aov, samples, os.fsencode(path)))

    def resolve_color(self, iteration: int, color_dev: int = 0) -> np.ndarray:
        """``image / iteration`` as Float3: the denoiser's colour input (main.cpp:194-201)."""
        out = np.empty((self.n_pixels, 3), np.float32)
        _check(self.lib.b2pt_resolve_color(self._h, iteration, color_dev or None, out.ctypes.data))
        return out

    def last_loop_ms(self) -> float:
        return float(self.lib.b2pt_last_loop_ms(self._h))

    def live_counts(self) -> np.ndarray:
        buf = np.zeros(64, np.int32)
        k = _check(self.lib.b2pt_live_counts(self._h, buf.ctypes.data_as(C.POINTER(C.c_int32)), 64))
        return buf[:k].copy()

    def walk_counts(self, long_walks: bool = False) -> np.ndarray:
        """Per-depth mesh-walk queue lengths of the last iteration (or the long-walk hand-offs)."""
        a, b = np.zeros(64, np.int32), np.zeros(64, np.int32)
        k = _check(self.lib.b2pt_walk_counts(self._h, a.ctypes.data_as(C.POINTER(C.c_int32)),
                                             b.ctypes.data_as(C.POINTER(C.c_int32)), 64))
        return (b if long_walks else a)[:k].copy()

    def launch_count(self) -> int:
        return int(self.lib.b2pt_launch_count(self._h))

    # -- zero-copy hand-off ----------------------------------------------------------------
    def device_image_ptr(self) -> int:
        return int(self.lib.b2pt_device_image(self._h) or 0)

    def device_albedo_ptr(self) -> int:
        return int(self.lib.b2pt_device_albedo(self._h) or 0)

    def set_device_image_ptr(self, ptr: int) -> None:
        _check(self.lib.b2pt_set_device_image(self._h, C.c_void_p(ptr)))

    def set_stream_ptr(self, ptr: int) -> None:
        _check(self.lib.b2pt_set_stream(self._h, C.c_void_p(ptr) if ptr else None))

    def profile_iteration(self, iteration: int) -> Dict[str, float]:
        ms = (C.c_float * 5)()
        _check(self.lib.b2pt_profile_iteration(self._h, iteration, ms))
        return dict(zip(["generate", "intersect", "sort", "shade", "iteration"], [float(x) for x in ms]))

    def profile_kernels(self, iteration: int) -> Dict[str, float]:
        """Per-kernel device time of one iteration (b2pt_profile_kernels), ms."""
        ms = (C.c_float * 8)()
        _check(self.lib.b2pt_profile_kernels(self._h, iteration, ms))
        names = ["generate", "analytic", "walk", "walk_long", "finish", "sort", "shade", "iteration"]
        return dict(zip(names, [float(x) for x in ms]))

    def stream_ptr(self) -> int:
        return int(self.lib.b2pt_stream(self._h) or 0)

    def tonemap_rgba8(self, dst_ptr: int, iteration: int, src_ptr: int = 0) -> None:
        _check(self.lib.b2pt_tonemap_rgba8(self._h, C.c_void_p(src_ptr) if src_ptr else None, iteration,
                                           C.c_void_p(dst_ptr)))

    # -- parity / introspection --------------------------------------------------------------
    def stage(self, depth: int, name: str) -> np.ndarray:
        sid, dt, cols = abi.STAGES[name]
        cap = self.n_pixels * cols
        buf = np.empty(cap, dt)
        got = self.lib.b2pt_stage_read(self._h, depth, sid, buf.ctypes.data, buf.nbytes)
        if got < 0:
            _check(int(got))
        n = got // 4 // cols
        out = buf[: n * cols].copy()
        return out.reshape(n, cols) if cols > 1 else out

    def stages(self) -> List[Dict[str, np.ndarray]]:
        """All recorded stage arrays of the last iteration (needs record_stages=1)."""
        res = []
        for d in range(max(self.pod.trace_depth, 1)):
            try:
                n = len(self.stage(d, "ray_pixel"))
            except B2ptError:
                break
            if n == 0:
                break
            res.append({name: self.stage(d, name) for name in abi.STAGES})
        return res

    def bvh_info(self, geom: int) -> abi.BvhInfo:
        info = abi.BvhInfo()
        _check(self.lib.b2pt_bvh_info(self._h, geom, C.byref(info)))
        return info


# ---------------------------------------------------------------------------------------
# Pipeline: pathtrace() one iteration per call, the next ones rendered ahead (csrc/pipe.cu)
# ---------------------------------------------------------------------------------------
class Pipeline:
    """``lanes`` contexts on one GPU behind the one-iteration-per-call contract of
    ``pathtrace(pbo, frame, iter)`` (apps/src/main.cpp:255, apps/src/pathtrace.cu:663-668)."""

    def __init__(self, scene, options: Optional[abi.Options] = None, lanes: int = 4, **opt_kw):
        self.lib = load_library()
        self.pod: PodScene = scene.pod if isinstance(scene, Scene) else scene
        self.options = options if options is not None else abi.default_options(**opt_kw)
        self._cscene = self.pod.as_ctypes()
        self._h = C.c_void_p()
        _check(self.lib.b2pt_pipe_create(C.byref(self._cscene), C.byref(self.options), lanes, C.byref(self._h)))
        self.n_pixels = self.pod.n_pixels
        self.lanes = lanes

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b2pt_pipe_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def pathtrace(self, iteration: int, image: np.ndarray, albedo: Optional[np.ndarray]) -> None:
        _check(self.lib.b2pt_pipe_pathtrace(self._h, iteration, image.ctypes.data,
                                            albedo.ctypes.data if albedo is not None else None))

    def reset(self, camera: Optional[np.ndarray] = None) -> None:
        if camera is None:
            _check(self.lib.b2pt_pipe_reset(self._h, None))
            return
        cam = abi.Camera()
        C.memmove(C.byref(cam), np.ascontiguousarray(camera).ctypes.data, C.sizeof(abi.Camera))
        _check(self.lib.b2pt_pipe_reset(self._h, C.byref(cam)))

    def device_image_ptr(self) -> int:
        return int(self.lib.b2pt_pipe_device_image(self._h) or 0)

    def device_albedo_ptr(self) -> int:
        return int(self.lib.b2pt_pipe_device_albedo(self._h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.b2pt_pipe_launch_count(self._h))

    def misses(self) -> int:
        return int(self.lib.b2pt_pipe_misses(self._h))

    def last_loop_ms(self) -> float:
        """Depth-loop time of the iteration the last call consumed; -1 until the call after the first query
        (the pipe only pays for the measurement once a host asks for it)."""
        return float(self.lib.b2pt_pipe_last_loop_ms(self._h))

    def tonemap_rgba8(self, dst_ptr: int, iteration: int, src_ptr: int = 0) -> None:
        """``sendImageToPBO`` of the running sum (or of ``src_ptr``); returns when ``dst_ptr`` is written."""
        _check(self.lib.b2pt_pipe_tonemap_rgba8(self._h, C.c_void_p(src_ptr) if src_ptr else None, iteration,
                                                C.c_void_p(dst_ptr)))


# ---------------------------------------------------------------------------------------
# several GPUs (csrc/multi.cu): one frame = one iteration per GPU, combined on the device
# ---------------------------------------------------------------------------------------
class MultiRenderer:
    """One process, several GPUs (``b2pt_multi_*``): ``pathtrace(first)`` renders iterations ``first ..
    first + G - 1`` at once, one per GPU, and returns the running sum including all of them -- bit-identical to
    one GPU rendering them one after the other.  ``devices`` may name a device twice (members share it)."""

    def __init__(self, scene, options: Optional[abi.Options] = None, devices=None, lanes: int = 4, **opt_kw):
        self.lib = load_library()
        self.pod: PodScene = scene.pod if isinstance(scene, Scene) else scene
        self.options = options if options is not None else abi.default_options(**opt_kw)
        self._cscene = self.pod.as_ctypes()
        self._h = C.c_void_p()
        if devices is None:
            devices = list(range(device_count()))
        arr = (C.c_int32 * len(devices))(*devices)
        _check(self.lib.b2pt_multi_create(C.byref(self._cscene), C.byref(self.options), len(devices), arr, lanes,
                                          C.byref(self._h)))
        self.members = len(devices)
        self.n_pixels = self.pod.n_pixels

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b2pt_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass

    def pathtrace(self, first_iteration: int, image: Optional[np.ndarray], albedo: Optional[np.ndarray] = None) -> None:
        _check(self.lib.b2pt_multi_pathtrace(self._h, first_iteration, image.ctypes.data if image is not None else None,
                                             albedo.ctypes.data if albedo is not None else None))

    def reset(self, camera: Optional[np.ndarray] = None) -> None:
        if camera is None:
            _check(self.lib.b2pt_multi_reset(self._h, None))
            return
        cam = abi.Camera()
        C.memmove(C.byref(cam), np.ascontiguousarray(camera).ctypes.data, C.sizeof(abi.Camera))
        _check(self.lib.b2pt_multi_reset(self._h, C.byref(cam)))

    def device_image_ptr(self) -> int:
        return int(self.lib.b2pt_multi_device_image(self._h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.b2pt_multi_launch_count(self._h))


class Shard:
    """One rank of a one-process-per-GPU job (``b2pt_shard_*``).  The cross-rank plumbing (exchange of the
    export blobs, the two barriers of a frame) belongs to the host: see :mod:`mygpuraytracer_b200.distributed`
    for the ``torch.distributed`` form."""

    def __init__(self, scene, options: Optional[abi.Options] = None, rank: int = 0, world: int = 1, lanes: int = 4, **opt_kw):
        self.lib = load_library()
        self.pod: PodScene = scene.pod if isinstance(scene, Scene) else scene
        self.options = options if options is not None else abi.default_options(**opt_kw)
        self._cscene = self.pod.as_ctypes()
        self._h = C.c_void_p()
        _check(self.lib.b2pt_shard_create(C.byref(self._cscene), C.byref(self.options), rank, world, lanes, C.byref(self._h)))
        self.rank, self.world, self.lanes = rank, world, lanes
        self.n_pixels = self.pod.n_pixels

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b2pt_shard_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass

    def export(self) -> bytes:
        n = int(self.lib.b2pt_shard_export_size(self._h))
        buf = C.create_string_buffer(n)
        _check(self.lib.b2pt_shard_export(self._h, buf, n))
        return buf.raw

    def connect(self, blobs) -> None:
        """``blobs``: the export blobs of all ranks in rank order (own entry included)."""
        per = len(blobs[0])
        if any(len(b) != per for b in blobs) or len(blobs) != self.world:
            raise ValueError("one export blob per rank, all of the same size")
        _check(self.lib.b2pt_shard_connect(self._h, b"".join(blobs), per))

    def stream_ptr(self) -> int:
        return int(self.lib.b2pt_shard_stream(self._h) or 0)

    def frame_begin(self, first_iteration: int) -> None:
        _check(self.lib.b2pt_shard_frame_begin(self._h, first_iteration))

    def frame_reduce(self) -> None:
        _check(self.lib.b2pt_shard_frame_reduce(self._h))

    def frame_image_ptr(self) -> int:
        return int(self.lib.b2pt_shard_frame_image(self._h) or 0)

    def frame_merge(self) -> None:
        _check(self.lib.b2pt_shard_frame_merge(self._h))

    def frame_end(self, image: Optional[np.ndarray] = None, albedo: Optional[np.ndarray] = None) -> None:
        _check(self.lib.b2pt_shard_frame_end(self._h, image.ctypes.data if image is not None else None,
                                             albedo.ctypes.data if albedo is not None else None))

    def reset(self, camera: Optional[np.ndarray] = None) -> None:
        if camera is None:
            _check(self.lib.b2pt_shard_reset(self._h, None))
            return
        cam = abi.Camera()
        C.memmove(C.byref(cam), np.ascontiguousarray(camera).ctypes.data, C.sizeof(abi.Camera))
        _check(self.lib.b2pt_shard_reset(self._h, C.byref(cam)))

    def sync(self) -> None:
        _check(self.lib.b2pt_shard_sync(self._h))

    def device_image_ptr(self) -> int:
        return int(self.lib.b2pt_shard_device_image(self._h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.b2pt_shard_launch_count(self._h))

    def misses(self) -> int:
        return int(self.lib.b2pt_shard_misses(self._h))

    def lane(self, k: int) -> "Renderer":
        """Lane k's context as a non-owning :class:`Renderer` (statistics, profiling)."""
        h = self.lib.b2pt_shard_lane(self._h, k)
        if not h:
            raise B2ptError(-6, "no such lane")
        return Renderer.view(h, self.pod)


# ---------------------------------------------------------------------------------------
# standalone primitives (apps/stream_compaction surface)
# ---------------------------------------------------------------------------------------
def scan_exclusive(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.int32)
    out = np.empty_like(a)
    _check(load_library().b2pt_scan_exclusive_i32(len(a), out.ctypes.data, a.ctypes.data))
    return out


def compact_nonzero(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.int32)
    out = np.empty_like(a)
    k = _check(load_library().b2pt_compact_nonzero_i32(len(a), out.ctypes.data, a.ctypes.data))
    return out[:k].copy()


def partition_perm(flags: np.ndarray):
    f = np.ascontiguousarray(flags, np.uint8)
    perm = np.empty(len(f), np.int32)
    k = _check(load_library().b2pt_partition_perm(len(f), f.ctypes.data, perm.ctypes.data))
    return perm, k


def sort_desc_perm(keys: np.ndarray) -> np.ndarray:
    k = np.ascontiguousarray(keys, np.int32)
    perm = np.empty(len(k), np.int32)
    _check(load_library().b2pt_sort_desc_perm(len(k), k.ctypes.data, perm.ctypes.data))
    return perm


def sort_material_ranks(material: np.ndarray, live: np.ndarray, n_materials: int, general: bool = False):
    """The renderer's material sort: (perm, rank, survivors); see b2pt_sort_material_ranks."""
    m = np.ascontiguousarray(material, np.uint8)
    f = np.ascontiguousarray(live, np.uint8)
    perm, rank = np.empty(len(m), np.int32), np.empty(len(m), np.int32)
    k = _check(load_library().b2pt_sort_material_ranks(len(m), m.ctypes.data, f.ctypes.data, int(n_materials), int(general),
                                                       perm.ctypes.data, rank.ctypes.data))
    return perm, rank, k


def radix_sort_pairs(keys: np.ndarray, vals: np.ndarray):
    k = np.ascontiguousarray(keys, np.uint32).copy()
    v = np.ascontiguousarray(vals, np.uint32).copy()
    _check(load_library().b2pt_radix_sort_pairs_u32(len(k), k.ctypes.data, v.ctypes.data))
    return k, v


# ---------------------------------------------------------------------------------------
# the reference's five free functions (global state, like apps/src/pathtrace.cu:118-128)
# ---------------------------------------------------------------------------------------
_hst_scene: Optional[Scene] = None
_renderer = None  # Renderer, or Pipeline after set_pipeline_lanes(n > 1)
_options_for_next_init: Optional[abi.Options] = None
_lanes_for_next_init: int = 1


class _Timer:
    """``timer()``: only the GPU stopwatch of the depth loop is reproduced."""

    def getGpuElapsedTimeForPreviousOperation(self) -> float:
        return _renderer.last_loop_ms() if _renderer is not None else 0.0


_timer = _Timer()


def timer() -> _Timer:
    return _timer


def set_options(options: Optional[abi.Options]) -> None:
    """Options the next :func:`pathtraceInit` uses (the reference's macros are compile time)."""
    global _options_for_next_init
    _options_for_next_init = options


def set_pipeline_lanes(lanes: int) -> None:
    """``lanes`` > 1: the next :func:`pathtraceInit` serves ``pathtrace`` from a :class:`Pipeline`."""
    global _lanes_for_next_init
    _lanes_for_next_init = max(1, int(lanes))


def pathtraceInit(scene: Scene) -> None:
    global _hst_scene, _renderer
    if _renderer is not None:
        _renderer.close()
    _hst_scene = scene
    opt = _options_for_next_init
    if opt is None:
        # scene.state.albedo is one array owned by the Scene and written only by pathtrace(), like the reference's
        # state.albedo vector: the unchanged AOV need not be copied again after iteration 1
        opt = abi.default_options(persistent_host_albedo=1)
    if _lanes_for_next_init > 1:
        _renderer = Pipeline(scene, opt, lanes=_lanes_for_next_init)
    else:
        _renderer = Renderer(scene, opt)


def pathtraceFree() -> None:
    global _renderer, _hst_scene
    if _renderer is not None:
        _renderer.close()
    _renderer = None
    _hst_scene = None


def pathtrace(pbo, frame: int, iteration: int) -> None:
    """One reference ``pathtrace(pbo, frame, iter)`` call: render, then D2H of
    ``image`` and ``albedo`` into the scene's RenderState.  ``pbo`` and
    ``frame`` are accepted and ignored, as in the reference's AI_DENOISE build."""
    if _renderer is None or _hst_scene is None:
        raise B2ptError(-5, "pathtrace() before pathtraceInit()")
    _renderer.pathtrace(iteration, _hst_scene.state.image, _hst_scene.state.albedo)


def sendToGPU(pbo_ptr: int, iteration: int) -> None:
    """``sendToGPU``: tone-map the accumulated image into a device RGBA8 buffer."""
    if _renderer is None:
        raise B2ptError(-5, "sendToGPU() before pathtraceInit()")
    _renderer.tonemap_rgba8(pbo_ptr, iteration)

"""ctypes mirror of ``include/b2pt.h`` (the C ABI of the path-tracing hot path).

Only declarations live here: the structures of ``b2pt.h`` field for field, the
enum values, and :func:`declare` which attaches argument / return types to a
loaded ``libb2pt.so``.  Nothing in this module computes anything.

Reference surface mirrored (see ``include/b2pt.h`` for the per-symbol
citations): ``apps/src/pathtrace.h:6-10`` and the PODs of
``apps/src/sceneStructs.h``.
"""
from __future__ import annotations

import ctypes as C

B2PT_OK = 0
ERR_NAMES = {0: "OK", -1: "INVALID", -2: "CUDA", -3: "NOMEM", -4: "IO", -5: "STATE", -6: "RANGE"}

SPHERE, CUBE, TRIANGLE, OBJ = 0, 1, 2, 3
TRIG_NATIVE, TRIG_PORTABLE = 0, 1
RNG_SLOT, RNG_PIXEL = 0, 1

# stage ids of b2pt_stage_read: name -> (id, numpy dtype, columns)
STAGES = {
    "ray_origin": (0, "f4", 3),
    "ray_dir": (1, "f4", 3),
    "ray_pixel": (2, "i4", 1),
    "hit_t": (3, "f4", 1),
    "hit_normal": (4, "f4", 3),
    "hit_uv": (5, "f4", 2),
    "hit_geom": (6, "i4", 1),
    "hit_face": (7, "i4", 1),
    "hit_material": (8, "i4", 1),
    "sort_perm": (9, "i4", 1),
    "shaded_color": (10, "f4", 3),
    "shaded_bounces": (11, "i4", 1),
    "shaded_origin": (12, "f4", 3),
    "shaded_dir": (13, "f4", 3),
    "partition_pixel": (14, "i4", 1),
}


class Material(C.Structure):
    _fields_ = [
        ("color", C.c_float * 3),
        ("specular_exponent", C.c_float),
        ("specular_color", C.c_float * 3),
        ("has_reflective", C.c_float),
        ("has_refractive", C.c_float),
        ("index_of_refraction", C.c_float),
        ("emittance", C.c_float),
    ]


class Texture(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("channels", C.c_int32),
        ("reserved", C.c_int32),
        ("texels", C.c_void_p),
    ]


class Geom(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("material_id", C.c_int32),
        ("transform", C.c_float * 16),
        ("inverse_transform", C.c_float * 16),
        ("inv_transpose", C.c_float * 16),
        ("face_begin", C.c_int32),
        ("face_count", C.c_int32),
        ("tex_kd", C.c_int32),
        ("tex_ks", C.c_int32),
        ("tex_bump", C.c_int32),
        ("tex_ke", C.c_int32),
    ]


class Camera(C.Structure):
    _fields_ = [
        ("resolution", C.c_int32 * 2),
        ("position", C.c_float * 3),
        ("look_at", C.c_float * 3),
        ("view", C.c_float * 3),
        ("up", C.c_float * 3),
        ("right", C.c_float * 3),
        ("fov", C.c_float * 2),
        ("pixel_length", C.c_float * 2),
    ]


class Scene(C.Structure):
    _fields_ = [
        ("n_geoms", C.c_int32),
        ("n_materials", C.c_int32),
        ("n_textures", C.c_int32),
        ("n_faces", C.c_int32),
        ("geoms", C.POINTER(Geom)),
        ("materials", C.POINTER(Material)),
        ("textures", C.POINTER(Texture)),
        ("face_pos", C.POINTER(C.c_float)),
        ("face_uv", C.POINTER(C.c_float)),
        ("camera", Camera),
        ("trace_depth", C.c_int32),
        ("iterations", C.c_int32),
        ("face_material", C.POINTER(C.c_int32)),
        ("material_textures", C.POINTER(C.c_int32)),
    ]


class Options(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("antialiasing", C.c_int32),
        ("depth_of_field", C.c_int32),
        ("lens_radius", C.c_float),
        ("focal_distance", C.c_float),
        ("sort_by_material", C.c_int32),
        ("cache_first_bounce", C.c_int32),
        ("trig_mode", C.c_int32),
        ("rng_mode", C.c_int32),
        ("use_bvh", C.c_int32),
        ("record_stages", C.c_int32),
        ("use_graph", C.c_int32),
        ("concurrent_contexts", C.c_int32),
        ("persistent_host_albedo", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


def default_options(**kw) -> Options:
    """The defaults of ``b2pt_default_options`` (the reference's macros,
    apps/src/pathtrace.cu:36-42,279-280), computed without the library so the
    oracle wrapper can use them on a box with no CUDA build."""
    o = Options()
    o.struct_size = C.sizeof(Options)
    o.device = 0
    o.antialiasing = 1
    o.depth_of_field = 0
    o.lens_radius = 0.8
    o.focal_distance = 11.0
    o.sort_by_material = 1
    o.cache_first_bounce = 0
    o.trig_mode = TRIG_NATIVE
    o.rng_mode = RNG_SLOT
    o.use_bvh = 1
    o.record_stages = 0
    o.use_graph = 1
    o.concurrent_contexts = 1
    o.persistent_host_albedo = 0
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(f"B2ptOptions has no field {k!r}")
        setattr(o, k, v)
    return o


AOV_IMAGE, AOV_ALBEDO = 0, 1


class BvhInfo(C.Structure):
    _fields_ = [
        ("n_faces", C.c_int32),
        ("n_nodes", C.c_int32),
        ("max_depth", C.c_int32),
        ("build_ms", C.c_float),
    ]


class LoadOverrides(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("iterations", C.c_int32),
        ("depth", C.c_int32),
        ("per_face_materials", C.c_int32),
    ]


# Every symbol include/b2pt.h declares: name -> (restype, argtypes).
_vp, _i32, _i64, _f32p = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_float)
_i32p, _u8p, _u32p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
SYMBOLS = {
    "b2pt_default_options": (None, [C.POINTER(Options)]),
    "b2pt_create": (C.c_int, [C.POINTER(Scene), C.POINTER(Options), C.POINTER(_vp)]),
    "b2pt_create_shared": (C.c_int, [_vp, C.POINTER(Scene), C.POINTER(Options), C.POINTER(_vp)]),
    "b2pt_destroy": (None, [_vp]),
    "b2pt_set_camera": (C.c_int, [_vp, C.POINTER(Camera)]),
    "b2pt_reset_accum": (C.c_int, [_vp]),
    "b2pt_render": (C.c_int, [_vp, _i32, _i32, _i32]),
    "b2pt_sync": (C.c_int, [_vp]),
    "b2pt_read_accum": (C.c_int, [_vp, _vp, _vp]),
    "b2pt_pathtrace": (C.c_int, [_vp, _i32, _vp, _vp]),
    "b2pt_pipe_create": (C.c_int, [C.POINTER(Scene), C.POINTER(Options), _i32, C.POINTER(_vp)]),
    "b2pt_pipe_destroy": (None, [_vp]),
    "b2pt_pipe_pathtrace": (C.c_int, [_vp, _i32, _vp, _vp]),
    "b2pt_pipe_reset": (C.c_int, [_vp, C.POINTER(Camera)]),
    "b2pt_pipe_device_image": (_vp, [_vp]),
    "b2pt_pipe_device_albedo": (_vp, [_vp]),
    "b2pt_pipe_tonemap_rgba8": (C.c_int, [_vp, _vp, _i32, _vp]),
    "b2pt_pipe_lanes": (_i32, [_vp]),
    "b2pt_pipe_lane": (_vp, [_vp, _i32]),
    "b2pt_pipe_launch_count": (_i64, [_vp]),
    "b2pt_pipe_misses": (_i64, [_vp]),
    "b2pt_pipe_last_loop_ms": (C.c_float, [_vp]),
    "b2pt_multi_create": (C.c_int, [C.POINTER(Scene), C.POINTER(Options), _i32, _i32p, _i32, C.POINTER(_vp)]),
    "b2pt_multi_destroy": (None, [_vp]),
    "b2pt_multi_pathtrace": (C.c_int, [_vp, _i32, _vp, _vp]),
    "b2pt_multi_reset": (C.c_int, [_vp, C.POINTER(Camera)]),
    "b2pt_multi_members": (_i32, [_vp]),
    "b2pt_multi_member": (_vp, [_vp, _i32]),
    "b2pt_multi_device_image": (_vp, [_vp]),
    "b2pt_multi_launch_count": (_i64, [_vp]),
    "b2pt_shard_create": (C.c_int, [C.POINTER(Scene), C.POINTER(Options), _i32, _i32, _i32, C.POINTER(_vp)]),
    "b2pt_shard_destroy": (None, [_vp]),
    "b2pt_shard_export_size": (_i64, [_vp]),
    "b2pt_shard_export": (C.c_int, [_vp, _vp, _i64]),
    "b2pt_shard_connect": (C.c_int, [_vp, _vp, _i64]),
    "b2pt_shard_frame_begin": (C.c_int, [_vp, _i32]),
    "b2pt_shard_stream": (_vp, [_vp]),
    "b2pt_shard_frame_reduce": (C.c_int, [_vp]),
    "b2pt_shard_frame_image": (_vp, [_vp]),
    "b2pt_shard_frame_merge": (C.c_int, [_vp]),
    "b2pt_shard_frame_end": (C.c_int, [_vp, _vp, _vp]),
    "b2pt_shard_reset": (C.c_int, [_vp, C.POINTER(Camera)]),
    "b2pt_shard_sync": (C.c_int, [_vp]),
    "b2pt_shard_device_image": (_vp, [_vp]),
    "b2pt_shard_device_slice": (_vp, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "b2pt_shard_lanes": (_i32, [_vp]),
    "b2pt_shard_lane": (_vp, [_vp, _i32]),
    "b2pt_shard_launch_count": (_i64, [_vp]),
    "b2pt_shard_misses": (_i64, [_vp]),
    "b2pt_device_image": (_vp, [_vp]),
    "b2pt_device_albedo": (_vp, [_vp]),
    "b2pt_set_device_image": (C.c_int, [_vp, _vp]),
    "b2pt_stream": (_vp, [_vp]),
    "b2pt_set_stream": (C.c_int, [_vp, _vp]),
    "b2pt_profile_iteration": (C.c_int, [_vp, _i32, _f32p]),
    "b2pt_profile_kernels": (C.c_int, [_vp, _i32, _f32p]),
    "b2pt_last_loop_ms": (C.c_float, [_vp]),
    "b2pt_tonemap_rgba8": (C.c_int, [_vp, _vp, _i32, _vp]),
    "b2pt_resolve_rgb8": (C.c_int, [_vp, _i32, _i32, _i32, _vp]),
    "b2pt_save_png": (C.c_int, [_vp, _i32, _i32, C.c_char_p]),
    "b2pt_save_hdr": (C.c_int, [_vp, _i32, _i32, C.c_char_p]),
    "b2pt_resolve_color": (C.c_int, [_vp, _i32, _vp, _vp]),
    "b2pt_live_counts": (C.c_int, [_vp, _i32p, _i32]),
    "b2pt_walk_counts": (C.c_int, [_vp, _i32p, _i32p, _i32]),
    "b2pt_launch_count": (_i64, [_vp]),
    "b2pt_stage_read": (_i64, [_vp, _i32, _i32, _vp, _i64]),
    "b2pt_bvh_info": (C.c_int, [_vp, _i32, C.POINTER(BvhInfo)]),
    "b2pt_scan_exclusive_i32": (C.c_int, [_i32, _vp, _vp]),
    "b2pt_compact_nonzero_i32": (C.c_int, [_i32, _vp, _vp]),
    "b2pt_partition_perm": (C.c_int, [_i32, _vp, _vp]),
    "b2pt_sort_desc_perm": (C.c_int, [_i32, _vp, _vp]),
    "b2pt_sort_material_ranks": (C.c_int, [_i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "b2pt_radix_sort_pairs_u32": (C.c_int, [_i32, _vp, _vp]),
    "b2pt_scene_load": (C.c_int, [C.c_char_p, C.POINTER(LoadOverrides), C.POINTER(_vp)]),
    "b2pt_scene_view": (C.POINTER(Scene), [_vp]),
    "b2pt_scene_image_name": (C.c_char_p, [_vp]),
    "b2pt_scene_warnings": (C.c_char_p, [_vp]),
    "b2pt_scene_free": (None, [_vp]),
    "b2pt_last_error": (C.c_char_p, []),
    "b2pt_abi_version": (C.c_int, []),
    "b2pt_device_count": (C.c_int, []),
}


def declare(lib: C.CDLL) -> C.CDLL:
    """Attach prototypes; raises AttributeError naming any missing export."""
    missing = []
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise AttributeError("libb2pt.so lacks symbols declared in include/b2pt.h: " + ", ".join(missing))
    return lib

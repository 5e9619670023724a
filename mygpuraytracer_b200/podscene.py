"""POD scenes: numpy-backed ``B2ptScene`` values and the ``.b2s`` container.

A :class:`PodScene` owns the arrays a ``B2ptScene`` (include/b2pt.h) points at:
geoms, materials, textures, the flat face arrays and the camera.  It can be
read from / written to a ``.b2s`` file, the tiny binary container the
reference-side tools under ``oracle/ref_driver`` emit after loading a scene
with the reference's own ``apps/src/scene.cpp`` -- this is how the parity
tests hand the SAME inputs to the reference, the oracle and the CUDA path.

Layout of ``.b2s`` (little endian)::

    char  magic[8] = "B2SCENE1"
    int32 n_geoms, n_materials, n_textures, n_faces, trace_depth, iterations
    B2ptCamera  camera
    B2ptGeom    geoms[n_geoms]
    B2ptMaterial materials[n_materials]
    per texture: int32 w, h, channels, reserved; uint8 texels[w*h*channels]
    float face_pos[n_faces*9]; float face_uv[n_faces*6]
    optional: char tag[4] = "PFM1"; int32 face_material[n_faces]; int32 material_textures[4*n_materials]
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import abi

GEOM_DTYPE = np.dtype(
    [
        ("type", "<i4"),
        ("material_id", "<i4"),
        ("transform", "<f4", (16,)),
        ("inverse_transform", "<f4", (16,)),
        ("inv_transpose", "<f4", (16,)),
        ("face_begin", "<i4"),
        ("face_count", "<i4"),
        ("tex_kd", "<i4"),
        ("tex_ks", "<i4"),
        ("tex_bump", "<i4"),
        ("tex_ke", "<i4"),
    ]
)
MATERIAL_DTYPE = np.dtype(
    [
        ("color", "<f4", (3,)),
        ("specular_exponent", "<f4"),
        ("specular_color", "<f4", (3,)),
        ("has_reflective", "<f4"),
        ("has_refractive", "<f4"),
        ("index_of_refraction", "<f4"),
        ("emittance", "<f4"),
    ]
)
CAMERA_DTYPE = np.dtype(
    [
        ("resolution", "<i4", (2,)),
        ("position", "<f4", (3,)),
        ("look_at", "<f4", (3,)),
        ("view", "<f4", (3,)),
        ("up", "<f4", (3,)),
        ("right", "<f4", (3,)),
        ("fov", "<f4", (2,)),
        ("pixel_length", "<f4", (2,)),
    ]
)
assert GEOM_DTYPE.itemsize == C.sizeof(abi.Geom)
assert MATERIAL_DTYPE.itemsize == C.sizeof(abi.Material) == 44
assert CAMERA_DTYPE.itemsize == C.sizeof(abi.Camera) == 84


@dataclass
class PodScene:
    geoms: np.ndarray  # GEOM_DTYPE[n_geoms]
    materials: np.ndarray  # MATERIAL_DTYPE[n_materials]
    camera: np.ndarray  # CAMERA_DTYPE scalar array, shape (1,)
    textures: List[np.ndarray] = field(default_factory=list)  # uint8 [h, w, c]
    face_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    face_uv: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    trace_depth: int = 8
    iterations: int = 1
    face_material: Optional[np.ndarray] = None      # int32[n_faces] or None (the reference: one material per OBJ)
    material_textures: Optional[np.ndarray] = None  # int32[n_materials, 4] (kd, ks, bump, ke) or None

    # -- convenience -----------------------------------------------------------
    @property
    def width(self) -> int:
        return int(self.camera["resolution"][0][0])

    @property
    def height(self) -> int:
        return int(self.camera["resolution"][0][1])

    @property
    def n_pixels(self) -> int:
        return self.width * self.height

    def copy(self) -> "PodScene":
        return PodScene(
            self.geoms.copy(), self.materials.copy(), self.camera.copy(), [t.copy() for t in self.textures],
            self.face_pos.copy(), self.face_uv.copy(), self.trace_depth, self.iterations,
            None if self.face_material is None else self.face_material.copy(),
            None if self.material_textures is None else self.material_textures.copy(),
        )

    # -- ctypes view -------------------------------------------------------------
    def as_ctypes(self) -> abi.Scene:
        """Build a ``B2ptScene`` pointing into this object's arrays.  The
        returned structure keeps references to every buffer it points at."""
        self.geoms = np.ascontiguousarray(self.geoms, GEOM_DTYPE)
        self.materials = np.ascontiguousarray(self.materials, MATERIAL_DTYPE)
        self.face_pos = np.ascontiguousarray(self.face_pos, np.float32).reshape(-1, 9)
        self.face_uv = np.ascontiguousarray(self.face_uv, np.float32).reshape(-1, 6)
        self.textures = [np.ascontiguousarray(t, np.uint8) for t in self.textures]
        s = abi.Scene()
        s.n_geoms = len(self.geoms)
        s.n_materials = len(self.materials)
        s.n_textures = len(self.textures)
        s.n_faces = len(self.face_pos)
        s.geoms = C.cast(self.geoms.ctypes.data, C.POINTER(abi.Geom))
        s.materials = C.cast(self.materials.ctypes.data, C.POINTER(abi.Material))
        tex = (abi.Texture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            h, w, c = t.shape
            tex[i].width, tex[i].height, tex[i].channels = w, h, c
            tex[i].texels = t.ctypes.data
        s.textures = C.cast(tex, C.POINTER(abi.Texture))
        s.face_pos = C.cast(self.face_pos.ctypes.data, C.POINTER(C.c_float))
        s.face_uv = C.cast(self.face_uv.ctypes.data, C.POINTER(C.c_float))
        C.memmove(C.byref(s.camera), self.camera.ctypes.data, C.sizeof(abi.Camera))
        s.trace_depth = int(self.trace_depth)
        s.iterations = int(self.iterations)
        if self.face_material is not None:
            self.face_material = np.ascontiguousarray(self.face_material, np.int32).reshape(-1)
            s.face_material = C.cast(self.face_material.ctypes.data, C.POINTER(C.c_int32))
        if self.material_textures is not None:
            self.material_textures = np.ascontiguousarray(self.material_textures, np.int32).reshape(-1, 4)
            s.material_textures = C.cast(self.material_textures.ctypes.data, C.POINTER(C.c_int32))
        s._keepalive = (self, tex)  # noqa: SLF001 - pin the backing storage
        return s

    # -- .b2s ------------------------------------------------------------------------
    @staticmethod
    def load(path: str) -> "PodScene":
        with open(path, "rb") as f:
            buf = f.read()
        if buf[:8] != b"B2SCENE1":
            raise ValueError(f"{path}: not a .b2s file")
        ng, nm, nt, nf, depth, iters = struct.unpack_from("<6i", buf, 8)
        off = 32
        cam = np.frombuffer(buf, CAMERA_DTYPE, 1, off).copy()
        off += CAMERA_DTYPE.itemsize
        geoms = np.frombuffer(buf, GEOM_DTYPE, ng, off).copy()
        off += GEOM_DTYPE.itemsize * ng
        mats = np.frombuffer(buf, MATERIAL_DTYPE, nm, off).copy()
        off += MATERIAL_DTYPE.itemsize * nm
        texs = []
        for _ in range(nt):
            w, h, c, _r = struct.unpack_from("<4i", buf, off)
            off += 16
            texs.append(np.frombuffer(buf, np.uint8, w * h * c, off).reshape(h, w, c).copy())
            off += w * h * c
        pos = np.frombuffer(buf, "<f4", nf * 9, off).reshape(nf, 9).copy()
        off += nf * 36
        uv = np.frombuffer(buf, "<f4", nf * 6, off).reshape(nf, 6).copy()
        off += nf * 24
        fm = mt = None
        if buf[off:off + 4] == b"PFM1":
            off += 4
            fm = np.frombuffer(buf, "<i4", nf, off).copy()
            off += 4 * nf
            mt = np.frombuffer(buf, "<i4", 4 * nm, off).reshape(nm, 4).copy()
            off += 16 * nm
        if off != len(buf):
            raise ValueError(f"{path}: trailing bytes ({len(buf) - off})")
        return PodScene(geoms, mats, cam, texs, pos, uv, depth, iters, fm, mt)

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(b"B2SCENE1")
            f.write(struct.pack("<6i", len(self.geoms), len(self.materials), len(self.textures), len(self.face_pos),
                                self.trace_depth, self.iterations))
            f.write(np.ascontiguousarray(self.camera, CAMERA_DTYPE).tobytes())
            f.write(np.ascontiguousarray(self.geoms, GEOM_DTYPE).tobytes())
            f.write(np.ascontiguousarray(self.materials, MATERIAL_DTYPE).tobytes())
            for t in self.textures:
                h, w, c = t.shape
                f.write(struct.pack("<4i", w, h, c, 0))
                f.write(np.ascontiguousarray(t, np.uint8).tobytes())
            f.write(np.ascontiguousarray(self.face_pos, "<f4").tobytes())
            f.write(np.ascontiguousarray(self.face_uv, "<f4").tobytes())
            if self.face_material is not None and self.material_textures is not None:
                f.write(b"PFM1")
                f.write(np.ascontiguousarray(self.face_material, "<i4").tobytes())
                f.write(np.ascontiguousarray(self.material_textures, "<i4").tobytes())

    @staticmethod
    def from_ctypes(s: abi.Scene) -> "PodScene":
        """Deep-copy a ``B2ptScene`` (e.g. the view of a loaded scene)."""
        geoms = np.ctypeslib.as_array(C.cast(s.geoms, C.POINTER(C.c_uint8)), (s.n_geoms * GEOM_DTYPE.itemsize,))
        geoms = geoms.view(GEOM_DTYPE).copy() if s.n_geoms else np.zeros(0, GEOM_DTYPE)
        mats = np.ctypeslib.as_array(C.cast(s.materials, C.POINTER(C.c_uint8)), (s.n_materials * 44,))
        mats = mats.view(MATERIAL_DTYPE).copy() if s.n_materials else np.zeros(0, MATERIAL_DTYPE)
        cam = np.frombuffer(bytes(s.camera), CAMERA_DTYPE, 1).copy()
        texs = []
        for i in range(s.n_textures):
            t = s.textures[i]
            n = t.width * t.height * t.channels
            a = np.ctypeslib.as_array(C.cast(t.texels, C.POINTER(C.c_uint8)), (n,)) if n else np.zeros(0, np.uint8)
            texs.append(a.reshape(t.height, t.width, t.channels).copy())
        nf = s.n_faces
        pos = np.ctypeslib.as_array(s.face_pos, (nf * 9,)).reshape(nf, 9).copy() if nf else np.zeros((0, 9), np.float32)
        uv = np.ctypeslib.as_array(s.face_uv, (nf * 6,)).reshape(nf, 6).copy() if nf else np.zeros((0, 6), np.float32)
        fm = np.ctypeslib.as_array(s.face_material, (nf,)).copy() if (s.face_material and nf) else None
        mt = (np.ctypeslib.as_array(s.material_textures, (s.n_materials * 4,)).reshape(-1, 4).copy()
              if s.material_textures else None)
        return PodScene(geoms, mats, cam, texs, pos, uv, s.trace_depth, s.iterations, fm, mt)

"""Run tree for the scene loader: scenes, the stand-in mesh, its MTL and textures.

The reference resolves everything relative to the working directory of the
executable (``../models/...`` in the scene file, ``../models/materials`` as the
MTL search path at apps/src/scene.cpp:41, ``..\\textures\\...`` inside the MTL).
:func:`prepare` lays the same tree out under a directory of your choice::

    <root>/bin/                      (working directory of the reference)
    <root>/scenes/<name>_<W>x<H>.txt
    <root>/models/Intergalactic_Spaceship-(Wavefront).obj    STAND-IN MESH
    <root>/models/materials/Intergalactic_Spaceship-(Wavefront).mtl
    <root>/textures/Intergalactic Spaceship_{color_4,rough,emi,nmap_2_Tris}.{jpg|ppm}

Textures: the reference ships seven 4096x4096 JPEGs.  They are data, are not
committed to this repository, and do not exist on the GPU box unless the build
step (``__graft_entry__.build``) copied them into the git-ignored
``assets/_gen/textures``.  When they are absent, deterministic procedural
textures of the same size are generated instead and every report says so.
"""
from __future__ import annotations

import os
import shutil
from typing import Optional

import numpy as np

from . import scenes, standin_mesh

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_ROOT = os.path.join(REPO, "assets", "_gen")
TEXTURE_STEMS = {
    "map_Bump": "Intergalactic Spaceship_nmap_2_Tris",
    "map_Kd": "Intergalactic Spaceship_color_4",
    "map_Ks": "Intergalactic Spaceship_rough",
    "map_Ke": "Intergalactic Spaceship_emi",
}


def _procedural(kind: str, size: int) -> np.ndarray:
    """Deterministic RGB8 stand-ins with the statistics that matter to the
    shader: the emission map is black except for sparse window-like patches,
    the normal map is a perturbed (128,128,255), kd/ks are mid-range noise."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    if kind == "map_Ke":
        lit = ((np.sin(x * 97.0) > 0.995) & (np.sin(y * 61.0) > 0.9))
        img = np.zeros((size, size, 3), np.uint8)
        img[lit] = (255, 220, 160)
        return img
    if kind == "map_Bump":
        nx = 128 + 40 * np.sin(x * 211.0) * np.cos(y * 173.0)
        ny = 128 + 40 * np.cos(x * 157.0) * np.sin(y * 199.0)
        return np.stack([nx, ny, np.full_like(nx, 235.0)], -1).astype(np.uint8)
    base = 110 + 70 * np.sin(x * 37.0 + 3 * np.sin(y * 11.0)) * np.cos(y * 29.0)
    if kind == "map_Ks":
        g = np.clip(base * 0.8, 0, 255)
        return np.stack([g, g, g], -1).astype(np.uint8)
    return np.stack([np.clip(base, 0, 255), np.clip(base * 0.9 + 10, 0, 255), np.clip(base * 0.8 + 25, 0, 255)], -1).astype(np.uint8)


def _write_ppm(path: str, img: np.ndarray) -> None:
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def textures_are_reference(root: str = DEFAULT_ROOT) -> bool:
    return all(os.path.exists(os.path.join(root, "textures", s + ".jpg")) for s in TEXTURE_STEMS.values())


def prepare(root: str = DEFAULT_ROOT, triangles: int = standin_mesh.HEADLINE_TRIANGLES,
            reference_textures: Optional[str] = None, procedural_size: int = 4096) -> str:
    """Create / refresh the run tree and return its root."""
    for d in ("bin", "scenes", "models/materials", "textures"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    tex_dir = os.path.join(root, "textures")
    if reference_textures and os.path.isdir(reference_textures):
        for stem in TEXTURE_STEMS.values():
            src = os.path.join(reference_textures, stem + ".jpg")
            dst = os.path.join(tex_dir, stem + ".jpg")
            if os.path.exists(src) and not os.path.exists(dst):
                shutil.copyfile(src, dst)
    use_jpg = textures_are_reference(root)
    if not use_jpg:
        for kind, stem in TEXTURE_STEMS.items():
            p = os.path.join(tex_dir, stem + ".ppm")
            if not os.path.exists(p):
                _write_ppm(p, _procedural(kind, procedural_size))
    ext = ".jpg" if use_jpg else ".ppm"
    # the MTL of the spaceship (apps/models/materials/...mtl): one material, four maps
    mtl = ["newmtl Material", "Ns 96.078431", "Ka 1.000000 1.000000 1.000000", "Kd 0.640000 0.640000 0.640000",
           "Ks 0.500000 0.500000 0.500000", "Ke 0.000000 0.000000 0.000000", "Ni 2.000000", "d 1.000000", "illum 2"]
    mtl += [f"{k} ../textures/{stem}{ext}" for k, stem in TEXTURE_STEMS.items()]
    with open(os.path.join(root, "models", "materials", standin_mesh.MTL_NAME), "w") as f:
        f.write("\n".join(mtl) + "\n")
    set_mesh(root, triangles)
    return root


def set_mesh(root: str, triangles: int) -> str:
    """Point models/<spaceship>.obj at the stand-in with ~`triangles` faces."""
    src = standin_mesh.ensure_obj(os.path.join(root, "models"), triangles)
    dst = os.path.join(root, "models", standin_mesh.OBJ_NAME)
    if os.path.lexists(dst):
        os.remove(dst)
    os.symlink(os.path.basename(src), dst)
    return dst


def scene_file(name: str, width: int, height: int, iterations: int = 5000, depth: int = 8, root: str = DEFAULT_ROOT) -> str:
    """Write <root>/scenes/<name>_<W>x<H>_d<depth>.txt and return its path."""
    path = os.path.join(root, "scenes", f"{name}_{width}x{height}_d{depth}.txt")
    scenes.write_scene(name, path, width=width, height=height, iterations=iterations, depth=depth)
    return path

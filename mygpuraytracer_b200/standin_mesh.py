"""Deterministic stand-in for the missing spaceship mesh.

``apps/models/Intergalactic_Spaceship-(Wavefront).obj`` is listed in the
reference's ``.MISSING_LARGE_BLOBS`` and exists nowhere, so ``cornellObj.txt``
and ``cornellSpaceship.txt`` cannot load as shipped.  This
module writes a replacement that satisfies everything the reference loader and
intersection code assume (apps/src/scene.cpp:38-234,
apps/src/intersections.h:207-282):

* ``mtllib Intergalactic_Spaceship-(Wavefront).mtl`` + ``usemtl Material``
  (the loader uses MTL material 0 unconditionally, scene.cpp:68,134);
* every vertex has ``v/vt/vn``; faces are pre-triangulated (loader independent);
* closed surface, counter-clockwise seen from outside (back faces are culled by
  glm::intersectRayTriangle, gtx/intersect.inl:51);
* texture coordinates strictly inside (0,1) (the texel index is never wrapped
  or clamped, intersections.h:227,270-272);
* a hull of roughly 4 x 2 x 5 units around the origin, so it sits inside the
  Cornell box at ``TRANS 1 3 3`` (apps/scenes/cornellSpaceship.txt:139-143).

Every coordinate is a float32 printed with nine significant digits, so any
correctly rounding decimal parser (and tinyobj's) recovers the same bits.
EVERY REPORT THAT USES THIS MESH MUST SAY "stand-in mesh" AND GIVE THE
TRIANGLE COUNT.
"""
from __future__ import annotations

import math
import os

import numpy as np

MTL_NAME = "Intergalactic_Spaceship-(Wavefront).mtl"
OBJ_NAME = "Intergalactic_Spaceship-(Wavefront).obj"
HEADLINE_TRIANGLES = 250_000


def grid_for(target_triangles: int):
    """(nu, nv) with 2*nu*nv + 2*nu close to the target, nv = 2*nu."""
    nu = max(4, int(round(math.sqrt(max(target_triangles, 32) / 4.0))))
    return nu, 2 * nu


def build(target_triangles: int = HEADLINE_TRIANGLES):
    """Return (positions[V,3], uvs[V,2], normals[V,3], faces[T,3]) as float32 / int32."""
    nu, nv = grid_for(target_triangles)
    i = np.arange(nu + 1, dtype=np.float64)  # column nu duplicates column 0 (uv seam)
    j = np.arange(nv + 1, dtype=np.float64)
    phi = 2.0 * np.pi * (i % nu) / nu
    s = 0.02 + 0.96 * j / nv  # along the hull, never reaching the poles
    PH, S = np.meshgrid(phi, s, indexing="xy")  # [nv+1, nu+1]
    profile = np.sin(np.pi * S) ** 0.6
    wing = 1.0 + 0.55 * np.abs(np.cos(PH)) ** 8 * np.sin(np.pi * np.clip((S - 0.25) / 0.6, 0, 1)) ** 2
    ripple = 1.0 + 0.03 * np.sin(9.0 * PH) * np.sin(14.0 * np.pi * S)
    fin = 1.0 + 0.35 * np.clip(np.sin(PH), 0, 1) ** 12 * np.clip((S - 0.6) / 0.3, 0, 1)
    x = 1.45 * profile * wing * ripple * np.cos(PH)
    y = 0.8 * profile * ripple * fin * np.sin(PH)
    z = 5.0 * (S - 0.5) + 0.0 * PH
    pos = np.stack([x, y, z], -1).reshape(-1, 3)
    u = 0.02 + 0.96 * (np.arange(nu + 1) / nu)
    v = 0.02 + 0.96 * (np.arange(nv + 1) / nv)
    UU, VV = np.meshgrid(u, v, indexing="xy")
    uv = np.stack([UU, VV], -1).reshape(-1, 2)
    # end caps: one centre vertex each
    c0 = np.array([[0.0, 0.0, pos[:nu + 1, 2].mean()]])
    c1 = np.array([[0.0, 0.0, pos[-(nu + 1):, 2].mean()]])
    pos = np.concatenate([pos, c0, c1])
    uv = np.concatenate([uv, [[0.5, 0.008]], [[0.5, 0.992]]])
    ic0, ic1 = len(pos) - 2, len(pos) - 1

    def vid(col, row):
        return row * (nu + 1) + col

    cols, rows = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    a = vid(cols, rows).ravel()
    b = vid(cols + 1, rows).ravel()
    c = vid(cols, rows + 1).ravel()
    d = vid(cols + 1, rows + 1).ravel()
    side = np.concatenate([np.stack([a, b, d], -1), np.stack([a, d, c], -1)])
    cap_cols = np.arange(nu)
    cap0 = np.stack([np.full(nu, ic0), vid(cap_cols + 1, 0), vid(cap_cols, 0)], -1)
    cap1 = np.stack([np.full(nu, ic1), vid(cap_cols, nv), vid(cap_cols + 1, nv)], -1)
    faces = np.concatenate([side, cap0, cap1]).astype(np.int64)

    pos32 = pos.astype(np.float32)
    # orient every triangle outwards (the hull is star-shaped about its axis)
    p0, p1, p2 = (pos32[faces[:, k]].astype(np.float64) for k in range(3))
    n = np.cross(p1 - p0, p2 - p0)
    cen = (p0 + p1 + p2) / 3.0
    radial = cen.copy()
    radial[:, 2] = 0.0
    is_cap = np.arange(len(faces)) >= len(side)
    radial[is_cap] = np.where(cen[is_cap, 2:3] < 0, [[0, 0, -1.0]], [[0, 0, 1.0]])
    flip = (n * radial).sum(-1) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]

    # smooth vertex normals (only written to the file; the hot path never reads them)
    p0, p1, p2 = (pos32[faces[:, k]].astype(np.float64) for k in range(3))
    fn = np.cross(p1 - p0, p2 - p0)
    vn = np.zeros_like(pos)
    for k in range(3):
        np.add.at(vn, faces[:, k], fn)
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    vn = np.where(ln > 0, vn / np.maximum(ln, 1e-30), [[0, 1.0, 0]])
    return pos32, uv.astype(np.float32), vn.astype(np.float32), faces.astype(np.int32)


def write_obj(path: str, target_triangles: int = HEADLINE_TRIANGLES) -> int:
    """Write the stand-in OBJ; returns the triangle count."""
    pos, uv, vn, faces = build(target_triangles)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    lines = [
        "# STAND-IN MESH generated by mygpuraytracer_b200.standin_mesh (the reference's",
        "# Intergalactic_Spaceship-(Wavefront).obj is missing from its checkout).",
        f"# triangles: {len(faces)}",
        f"mtllib {MTL_NAME}",
        "o standin_hull",
    ]
    lines += ["v %.9g %.9g %.9g" % tuple(p) for p in pos.tolist()]
    lines += ["vt %.9g %.9g" % tuple(t) for t in uv.tolist()]
    lines += ["vn %.9g %.9g %.9g" % tuple(n) for n in vn.tolist()]
    lines += ["usemtl Material", "s off"]
    f1 = faces + 1
    lines += ["f %d/%d/%d %d/%d/%d %d/%d/%d" % (a, a, a, b, b, b, c, c, c) for a, b, c in f1.tolist()]
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        f.write("\n".join(lines))
        f.write("\n")
    os.replace(tmp, path)
    return len(faces)


def ensure_obj(directory: str, target_triangles: int = HEADLINE_TRIANGLES) -> str:
    """Write ``<directory>/standin_<T>.obj`` if absent and return its path."""
    path = os.path.join(directory, f"standin_{target_triangles}.obj")
    if not os.path.exists(path):
        write_obj(path, target_triangles)
    return path


if __name__ == "__main__":
    import sys

    out = sys.argv[1] if len(sys.argv) > 1 else OBJ_NAME
    tri = int(sys.argv[2]) if len(sys.argv) > 2 else HEADLINE_TRIANGLES
    print(write_obj(out, tri), "triangles ->", out)

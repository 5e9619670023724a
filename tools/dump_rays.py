"""Dump the rays of one iteration of cornellSpaceship that cross the mesh's bounding box, in the mesh's object space,
together with the distance limit the walk starts with (CPU experiment input for tools/exp_wide_leaf.c; not product).

    python tools/dump_rays.py OUT_DIR [width height triangles]

Writes OUT_DIR/tris.bin (n x 9 float32) and OUT_DIR/rays.bin (m x 8 float32: origin, direction, t_limit, depth).
The rays come from the CPU oracle's stage dumps, i.e. they are the rays the GPU walk kernel sees.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import abi, api, assets  # noqa: E402
from oracle import oracle  # noqa: E402

out = sys.argv[1]
w, h, tris = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (192, 108, 250000)
os.makedirs(out, exist_ok=True)
root = assets.prepare(os.path.join(out, "run"), triangles=tris, procedural_size=256)
pod = api.Scene(assets.scene_file("cornellSpaceship", w, h, root=root)).pod
g = int(np.nonzero(pod.geoms["type"] == abi.OBJ)[0][0])
G = pod.geoms[g]
fb, fc = int(G["face_begin"]), int(G["face_count"])
pos = np.ascontiguousarray(pod.face_pos[fb:fb + fc], np.float32).reshape(fc, 9)
pos.tofile(os.path.join(out, "tris.bin"))
inv = np.asarray(G["inverse_transform"], np.float32).reshape(4, 4).T  # column-major -> M[row, col]
lo, hi = pos.reshape(-1, 3).min(0) - 1e-3, pos.reshape(-1, 3).max(0) + 1e-3
img = np.zeros((pod.n_pixels, 3), np.float32)
rows = []
for it in (1, 2):
    stages = oracle.iteration_with_stages(pod, abi.default_options(), it, img, None)
    for d, st in enumerate(stages):
        o = st["ray_origin"] @ inv[:3, :3].T + inv[:3, 3]
        dd = st["ray_dir"] @ inv[:3, :3].T
        dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0, t1 = (lo - o) / dd, (hi - o) / dd
        tn = np.nanmax(np.minimum(t0, t1), axis=1)
        tf = np.nanmin(np.maximum(t0, t1), axis=1)
        # the limit the walk starts with: the closest analytic hit (a mesh winner has none closer than itself)
        t = st["hit_t"].astype(np.float32)
        mesh_won = st["hit_geom"] == g
        tlim = np.where((t > 0) & ~mesh_won, t * 1.0001 + 1e-5, np.inf).astype(np.float32)
        keep = (tn <= tf) & (tf >= 0) & (tn <= tlim)
        r = np.concatenate([o, dd, tlim[:, None], np.full((len(o), 1), d, np.float32)], 1).astype(np.float32)[keep]
        rows.append(r)
        print(f"iter {it} depth {d}: {len(o)} rays, {int(keep.sum())} cross the mesh box, {int(mesh_won.sum())} hit the mesh", flush=True)
rays = np.concatenate(rows)
rays.tofile(os.path.join(out, "rays.bin"))
print(len(rays), "rays ->", os.path.join(out, "rays.bin"))

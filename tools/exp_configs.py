"""The five BASELINE.json configurations on one GPU: ms/iteration and Mpaths/s (graph replay, one context)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import api, abi, assets
root = assets.prepare()
assets.set_mesh(root, 250000)
CASES = [
    ("config 1: cornell 800x800 depth 8", "cornell", 800, 800, 8, {}, 2000),
    ("config 2: cornellGlass 800x800 depth 8, DOF + AA, 5000 spp", "cornellGlass", 800, 800, 8, {"depth_of_field": 1}, 5000),
    ("config 3: cornellObj 1920x1080 depth 8 (LBVH build + traversal)", "cornellObj", 1920, 1080, 8, {}, 300),
    ("config 4: cornellSpaceship 1920x1080 depth 8", "cornellSpaceship", 1920, 1080, 8, {}, 300),
    ("config 5: cornellSpaceship 3840x2160 depth 12", "cornellSpaceship", 3840, 2160, 12, {}, 100),
]
out = []
for name, scene, w, h, depth, kw, iters in CASES:
    sc = api.Scene(assets.scene_file(scene, w, h, depth=depth, root=root))
    with api.Renderer(sc, abi.default_options(**kw)) as r:
        r.render(1, 10, 1); r.sync()
        t0 = time.time(); r.render(11, iters, 1); r.sync(); dt = time.time() - t0
        live = [int(x) for x in r.live_counts()]
        bvh = None
        try:
            import numpy as np
            g = int(np.nonzero(sc.pod.geoms["type"] == abi.OBJ)[0][0])
            b = r.bvh_info(g); bvh = {"triangles": int(b.n_faces), "build_ms": round(float(b.build_ms), 2), "max_depth": int(b.max_depth)}
        except Exception:
            pass
    row = {"config": name, "iterations": iters, "ms_per_iteration": round(dt / iters * 1e3, 4),
           "mpaths_per_s": round(w * h * iters / dt / 1e6, 1), "seconds_total": round(dt, 2), "segments_per_iteration": sum(live[:depth]), "bvh": bvh}
    out.append(row)
    print(json.dumps(row), flush=True)

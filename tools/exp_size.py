"""Time a few iterations of cornellSpaceship at a given resolution / mesh size (debug helper)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import api, abi, assets
w, h, tris, iters = (int(x) for x in sys.argv[1:5])
root = assets.prepare()
assets.set_mesh(root, tris)
sc = api.Scene(assets.scene_file("cornellSpaceship", w, h, root=root))
with api.Renderer(sc, abi.default_options(use_graph=0)) as r:
    t0 = time.time(); r.render(1, iters, 1); r.sync(); dt = time.time() - t0
    print(f"{w}x{h} tris={tris}: {dt / iters * 1e3:.3f} ms/iter, walks={int(r.walk_counts().sum())}", flush=True)

import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from mygpuraytracer_b200 import api, abi, assets
root = assets.prepare()
def run(scene, tris=250000, w=1920, h=1080, iters=30, **kw):
    if tris: assets.set_mesh(root, tris)
    sc = api.Scene(assets.scene_file(scene, w, h, root=root))
    with api.Renderer(sc, abi.default_options(**kw)) as r:
        r.render(1, 5, 1); r.sync()
        t0 = time.time(); r.render(6, iters, 1); r.sync(); dt = (time.time()-t0)/iters*1e3
        prof = r.profile_iteration(100)
        live = r.live_counts()
    print(f"{scene:18s} tris={tris:7d} {kw} ms/iter={dt:.3f} prof={ {k: round(v,3) for k,v in prof.items()} } live={list(live[:9])}", flush=True)
run("cornellGlass", 0)
run("cornellSpaceship", 250000)
run("cornellSpaceship", 50000)
run("cornellSpaceship", 1000)
os.environ["B2PT_TRAVERSAL_STATS"]="1"
run("cornellSpaceship", 250000, iters=1)

"""Render a few iterations of one configuration (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import api, abi, assets
tris = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cc = int(sys.argv[3]) if len(sys.argv) > 3 else 1  # 4: the shared-SM grids and unfused kernels of the timed region
root = assets.prepare()
assets.set_mesh(root, tris)
sc = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, root=root))
with api.Renderer(sc, abi.default_options(use_graph=0, concurrent_contexts=cc)) as r:
    r.render(1, iters, 1)
    r.sync()
    print(r.live_counts())

"""Small end-to-end case for compute-sanitizer: mesh scene, BVH build, 3 iterations, long walks, output hand-off."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mygpuraytracer_b200 import api, abi, assets
tmp = tempfile.mkdtemp()
root = assets.prepare(os.path.join(tmp, "run"), triangles=20000, procedural_size=256)
pod = api.Scene(assets.scene_file("cornellSpaceship", 160, 90, root=root)).pod
os.environ["B2PT_LONG_WALK"] = "6"   # force many hand-offs to the cooperative kernel
for kw in ({}, {"concurrent_contexts": 4}, {"antialiasing": 0, "cache_first_bounce": 1}, {"use_graph": 0, "sort_by_material": 0}):
    with api.Renderer(pod, abi.default_options(**kw)) as r:
        r.render(1, 3, 1)
        img, alb = r.read()
        r.resolve_rgb8(abi.AOV_IMAGE, 3)
        r.resolve_color(3)
        print(kw, float(img.sum()), int(r.walk_counts().sum()), int(r.walk_counts(True).sum()), flush=True)

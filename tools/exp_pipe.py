"""A few pipelined pathtrace() calls on the headline scene (for ncu captures of k_pipe_merge)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import api, abi, assets
root = assets.prepare()
assets.set_mesh(root, int(sys.argv[1]) if len(sys.argv) > 1 else 250000)
sc = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, root=root))
n = sc.pod.n_pixels
img, alb = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
with api.Pipeline(sc, abi.default_options(), lanes=4) as pipe:
    for it in range(1, (int(sys.argv[2]) if len(sys.argv) > 2 else 6) + 1):
        pipe.pathtrace(it, img, alb)
print("checksum", float(np.float64(img.sum())), "launches", 0)

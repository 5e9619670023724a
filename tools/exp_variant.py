"""One line per configuration: 1-stream and 3-stream ms/iteration + per-kernel times (headline scene)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import api, abi, assets
root = assets.prepare()
assets.set_mesh(root, 250000)
sc = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, root=root))
tag = sys.argv[1] if len(sys.argv) > 1 else ""
with api.Renderer(sc, abi.default_options()) as r:
    r.render(1, 5, 1); r.sync()
    t0 = time.time(); r.render(6, 30, 1); r.sync(); dt = (time.time() - t0) / 30 * 1e3
    prof = r.profile_kernels(100)
    nl = int(r.walk_counts(True).sum())
K = 3
rs = [api.Renderer(sc, abi.default_options()) for _ in range(K)]
for k, r in enumerate(rs): r.render(k + 1, 4, K)
for r in rs: r.sync()
iters = 60
t0 = time.time()
for k, r in enumerate(rs): r.render(100 + k, iters // K, K)
for r in rs: r.sync()
dt3 = (time.time() - t0) / iters * 1e3
for r in rs: r.close()
print(f"{tag:28s} 1s {dt:.3f} 3s {dt3:.3f} long={nl} " + " ".join(f"{k[:5]}={v:.3f}" for k, v in prof.items()), flush=True)

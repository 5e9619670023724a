"""Throughput with K concurrent contexts for the current B2PT_*_CTAS environment (headline scene)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import api, abi, assets
root = assets.prepare()
assets.set_mesh(root, 250000)
sc = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, root=root))
tag = sys.argv[1]
out = []
for K in [int(x) for x in sys.argv[2:]]:
    rs = [api.Renderer(sc, abi.default_options()) for _ in range(K)]
    for k, r in enumerate(rs): r.render(k + 1, 4, K)
    for r in rs: r.sync()
    iters = 60 // K * K
    t0 = time.time()
    for k, r in enumerate(rs): r.render(100 + k, iters // K, K)
    for r in rs: r.sync()
    out.append(f"K={K}: {(time.time() - t0) / iters * 1e3:.3f}")
    for r in rs: r.close()
print(f"{tag:24s} " + "  ".join(out), flush=True)

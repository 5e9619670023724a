"""Throughput with K contexts (streams) on one GPU rendering interleaved iterations."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import api, abi, assets
root = assets.prepare()
assets.set_mesh(root, 250000)
sc = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, root=root))
for K in (1, 2, 3, 4, 6):
    rs = [api.Renderer(sc, abi.default_options()) for _ in range(K)]
    for k, r in enumerate(rs): r.render(k + 1, 4, K)
    for r in rs: r.sync()
    iters = 48
    t0 = time.time()
    for k, r in enumerate(rs): r.render(100 + k, iters // K, K)
    for r in rs: r.sync()
    dt = time.time() - t0
    print(f"K={K}: {dt / iters * 1e3:.3f} ms/iteration  {1920*1080*iters/dt/1e6:.1f} Mpaths/s", flush=True)
    for r in rs: r.close()

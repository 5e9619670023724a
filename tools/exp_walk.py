"""Headline scene, one context: ms/iteration (graph replay, full-width grids and shared-SM grids), per-kernel times, walk
statistics.  `python tools/exp_walk.py [tag]`; B2PT_LIB selects another build of the library."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import abi, api, assets  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("B2PT_LIB", "default"))
root = assets.prepare()
assets.set_mesh(root, 250000)
sc = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, root=root))


def one(cc):
    with api.Renderer(sc, abi.default_options(concurrent_contexts=cc)) as r:
        r.render(1, 5, 1)
        r.sync()
        t0 = time.time()
        r.render(6, 40, 1)
        r.sync()
        dt = (time.time() - t0) / 40 * 1e3
        prof = [r.profile_kernels(100 + i) for i in range(5)]
        prof = {k: sorted(p[k] for p in prof)[2] for k in prof[0]}
        walks, longs = int(r.walk_counts().sum()), int(r.walk_counts(True).sum())
    return dt, prof, walks, longs


for cc in (1, 4):
    dt, prof, walks, longs = one(cc)
    print(f"{tag:24s} grids={'full' if cc == 1 else 'shared'} graph {dt:.3f} ms/iter  walks={walks} long={longs}  "
          + " ".join(f"{k[:6]}={v:.3f}" for k, v in prof.items()), flush=True)
# four contexts sharing one scene, interleaved iterations (the regime of bench.py's timed region)
K = 4
opt = abi.default_options(concurrent_contexts=K)
rs = [api.Renderer(sc, opt)]
rs += [api.Renderer(sc, opt, share=rs[0]) for _ in range(K - 1)]
for k, r in enumerate(rs):
    r.render(k + 1, 4, K)
for r in rs:
    r.sync()
iters = 160
t0 = time.time()
for k, r in enumerate(rs):
    r.render(100 + k, iters // K, K)
for r in rs:
    r.sync()
print(f"{tag:24s} 4 contexts: {(time.time() - t0) / iters * 1e3:.3f} ms/iter", flush=True)
for r in rs:
    r.close()

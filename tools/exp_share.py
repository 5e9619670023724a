"""Aggregate throughput of K contexts sharing one scene copy (the regime of bench.py's timed region) for a list of
grid configurations: `python tools/exp_share.py "K=4" "K=4,WALK=2" "K=5,ANALYTIC=3,SHADE=6" ...`.
Keys: K (contexts), WALK / LONG / ANALYTIC / FINISH / SHADE (resident CTAs per SM -> B2PT_*_CTAS), FUSE, LONGWALK (hand-off steps)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import abi, api, assets  # noqa: E402

root = assets.prepare()
assets.set_mesh(root, 250000)
sc = api.Scene(assets.scene_file("cornellSpaceship", 1920, 1080, root=root))
ENV = {"WALK": "B2PT_WALK_CTAS", "LONG": "B2PT_LONG_CTAS", "ANALYTIC": "B2PT_ANALYTIC_CTAS", "FINISH": "B2PT_FINISH_CTAS",
       "SHADE": "B2PT_SHADE_CTAS", "FUSE": "B2PT_FUSE", "LONGWALK": "B2PT_LONG_WALK"}
for spec in sys.argv[1:]:
    kv = dict(x.split("=") for x in spec.split(",") if x)
    K = int(kv.pop("K", 4))
    for e in ENV.values():
        os.environ.pop(e, None)
    for k, v in kv.items():
        os.environ[ENV[k]] = v
    opt = abi.default_options(concurrent_contexts=K)
    rs = [api.Renderer(sc, opt)]
    rs += [api.Renderer(sc, opt, share=rs[0]) for _ in range(K - 1)]
    for k, r in enumerate(rs):
        r.render(k + 1, 4, K)
    for r in rs:
        r.sync()
    best = 1e9
    for rep in range(3):
        iters = 40 * K
        t0 = time.time()
        for k, r in enumerate(rs):
            r.render(100 + 1000 * rep + k, iters // K, K)
        for r in rs:
            r.sync()
        best = min(best, (time.time() - t0) / iters * 1e3)
    for r in rs:
        r.close()
    print(f"{spec:40s} {best:.4f} ms/iter", flush=True)

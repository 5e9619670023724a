"""CPU model of k_intersect_analytic's exact-test rounds on the renderer's own rays (oracle stage dumps): how many
exact cube / sphere tests a ray runs and how many rounds a warp of 32 consecutive slots needs, for the index-order
loop of round 1 and for nearest-entry-first variants.  Experiment input, not product.
    python tools/exp_analytic_model.py [width height]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import abi, api, assets  # noqa: E402
from oracle import oracle  # noqa: E402

w, h = (int(x) for x in sys.argv[1:3]) if len(sys.argv) >= 3 else (384, 216)
root = assets.prepare("/tmp/model/run", triangles=1000, procedural_size=256)
pod = api.Scene(assets.scene_file("cornellSpaceship", w, h, root=root)).pod
G = pod.geoms
ng = len(G)
M = [np.asarray(g["transform"], np.float64).reshape(4, 4).T for g in G]
MI = [np.asarray(g["inverse_transform"], np.float64).reshape(4, 4).T for g in G]
typ = [int(g["type"]) for g in G]
print("types", typ)

def world_box(m):
    c = np.array([[x, y, z, 1.0] for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5)])
    wpts = c @ m.T
    lo, hi = wpts[:, :3].min(0), wpts[:, :3].max(0)
    ext = max(np.abs(lo).max(), np.abs(hi).max())
    scale = max(np.linalg.norm(m[:3, c]) for c in range(3))
    pad = 1e-3 + 1e-4 * ext
    return lo - pad, hi + pad, 2e-3 + 1.5e-4 * scale
boxes = [world_box(m) for m in M]

def slab(lo, hi, o, d):
    with np.errstate(divide="ignore", invalid="ignore"):
        i = 1.0 / d
        t0, t1 = (lo - o) * i, (hi - o) * i
    tn = np.nanmax(np.minimum(t0, t1), axis=1)
    tf = np.nanmin(np.maximum(t0, t1), axis=1)
    return tn, tf

def exact(g, o, d):
    """world-space hit distance of geom g (float64 restatement; -1 = miss)"""
    qo = o @ MI[g][:3, :3].T + MI[g][:3, 3]
    qd = d @ MI[g][:3, :3].T
    qd /= np.linalg.norm(qd, axis=1, keepdims=True)
    n = len(o)
    if typ[g] == abi.CUBE:
        tmin, tmax = np.full(n, -1e38), np.full(n, 1e38)
        with np.errstate(divide="ignore", invalid="ignore"):
            for a in range(3):
                ok = np.abs(qd[:, a]) > 1e-5
                t1, t2 = (-.5 - qo[:, a]) / qd[:, a], (.5 - qo[:, a]) / qd[:, a]
                ta, tb = np.minimum(t1, t2), np.maximum(t1, t2)
                u = ok & (ta > 0) & (ta > tmin)
                tmin = np.where(u, ta, tmin)
                tmax = np.where(ok & (tb < tmax), tb, tmax)
        hit = (tmax >= tmin) & (tmax > 0)
        t = np.where(tmin <= 0, tmax, tmin)
    else:
        vdd = (qo * qd).sum(1)
        rad = vdd * vdd - ((qo * qo).sum(1) - .25)
        sq = np.sqrt(np.maximum(rad, 0))
        t1, t2 = -vdd + sq, -vdd - sq
        hit = (rad >= 0) & ~((t1 < 0) & (t2 < 0))
        t = np.where((t1 > 0) & (t2 > 0), np.minimum(t1, t2), np.maximum(t1, t2))
    p = qo + qd * (t - 1e-4)[:, None]
    wp = p @ M[g][:3, :3].T + M[g][:3, 3]
    return np.where(hit, np.linalg.norm(o - wp, axis=1), -1.0)

img = np.zeros((pod.n_pixels, 3), np.float32)
stages = oracle.iteration_with_stages(pod, abi.default_options(), 1, img, None)
ana = [g for g in range(ng) if typ[g] in (abi.CUBE, abi.SPHERE)]
for dep, st in enumerate(stages[:4]):
    o, d = st["ray_origin"].astype(np.float64), st["ray_dir"].astype(np.float64)
    n = len(o)
    tn = np.full((n, ng), np.inf); cand = np.zeros((n, ng), bool); te = np.full((n, ng), -1.0)
    for g in ana:
        a, b = slab(boxes[g][0], boxes[g][1], o, d)
        cand[:, g] = (a <= b) & (b >= 0)
        tn[:, g] = np.where(cand[:, g], a, np.inf)
        te[:, g] = exact(g, o, d)
    slack = np.array([boxes[g][2] if g in ana else 0 for g in range(ng)])
    def run(order):  # order: per-ray candidate visiting order (n x k array of geom ids, -1 pad)
        tmin = np.full(n, np.inf); tests = np.zeros(n, int); rounds = np.zeros((n, order.shape[1]), bool)
        for k in range(order.shape[1]):
            g = order[:, k]; ok = g >= 0
            gi = np.where(ok, g, 0); r = np.arange(n)
            viable = ok & ~((tmin < np.inf) & (tn[r, gi] > tmin * 1.0001 + slack[gi]))
            t = te[r, gi]
            better = viable & (t > 0) & (t < tmin)
            tmin = np.where(better, t, tmin)
            tests += viable; rounds[:, k] = viable
        return tests, rounds
    # index order
    idx = np.where(cand, np.arange(ng)[None, :], 99)
    o_idx = np.sort(idx, 1); o_idx[o_idx == 99] = -1
    key = np.where(cand, tn, np.inf); o_near = np.argsort(key, 1, kind="stable").astype(int)
    o_near[np.take_along_axis(key, o_near, 1) == np.inf] = -1
    nw = (n + 31) // 32
    def warp_rounds_loop(rounds):  # the lane-own loop: one body execution per loop iteration in which any lane is viable
        pad = np.zeros((nw * 32, rounds.shape[1]), bool); pad[:n] = rounds
        return pad.reshape(nw, 32, -1).any(1).sum(1)
    def warp_rounds_skip(tests):  # lanes skip ahead to their next viable candidate: max tests over the lanes
        pad = np.zeros(nw * 32, int); pad[:n] = tests
        return pad.reshape(nw, 32).max(1)
    def warp_rounds_pool(tests):  # round 1 lane-own, the rest pooled over the warp
        pad = np.zeros(nw * 32, int); pad[:n] = tests
        p = pad.reshape(nw, 32)
        return (p > 0).any(1) + (np.maximum(p - 1, 0).sum(1) + 31) // 32
    print(f"depth {dep}: {n} rays, candidates/ray {cand.sum()/n:.2f}")
    for name, od in (("index order", o_idx), ("nearest first", o_near)):
        tests, rounds = run(od)
        hist = np.bincount(tests, minlength=5)[:6] / n
        print(f"  {name:14s} exact tests/ray {tests.mean():.2f} hist {np.round(hist,3)}  rounds/warp: loop {warp_rounds_loop(rounds).mean():.2f}"
              f"  skip {warp_rounds_skip(tests).mean():.2f}  pooled {warp_rounds_pool(tests).mean():.2f}")
    # type split of round 1 under nearest first
    g1 = o_near[:, 0]; isb = np.array([typ[g] == abi.CUBE if g >= 0 else False for g in g1]); iss = (g1 >= 0) & ~isb
    pad = np.zeros(nw * 32, int); pad[:n] = isb + 2 * iss
    p = pad.reshape(nw, 32)
    print(f"  round 1 (nearest first): warps with a box job {((p == 1).any(1)).mean():.2f}, with a sphere job {((p == 2).any(1)).mean():.2f}")

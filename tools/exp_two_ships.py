"""One mesh against two: cornellSpaceship and twoShips (a second OBJ geom, scaled by 1.5) at
1920x1080 with the 250 500-triangle stand-in mesh: walks, long walks and device time of the walk kernels per walk.
A ray that crosses both meshes' boxes walks both (one walk each), so the unit is the WALK, not the ray."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mygpuraytracer_b200 import abi, api, assets  # noqa: E402

root = assets.prepare()
assets.set_mesh(root, 250000)
for name in ("cornellSpaceship", "twoShips"):
    sc = api.Scene(assets.scene_file(name, 1920, 1080, root=root))
    for cc in (1, 4):
        with api.Renderer(sc, abi.default_options(concurrent_contexts=cc)) as r:
            r.render(1, 5, 1)
            r.sync()
            prof = [r.profile_kernels(100 + i) for i in range(5)]
            prof = {k: sorted(p[k] for p in prof)[2] for k in prof[0]}
            walks, longs = int(r.walk_counts().sum()), int(r.walk_counts(True).sum())
            segs = int(r.live_counts()[:8].sum())
        w = prof["walk"] + prof["walk_long"]
        print(f"{name:18s} grids={'full' if cc == 1 else 'shared'} segments={segs} queued rays={walks} handed off={longs} ({100 * longs / max(walks, 1):.1f} %) "
              f"walk={prof['walk']:.3f} long={prof['walk_long']:.3f} ms -> {1e6 * w / max(walks, 1):.2f} ns per queued ray; iteration {prof['iteration']:.3f} ms",
              flush=True)

"""Development helper (run under gpurun): step a big scene on the device only and print phase timings / counts."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from physecs_b200 import scenes as S  # noqa: E402
from physecs_b200.capi import Context  # noqa: E402

if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "terrain"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    t0 = time.time()
    if kind == "terrain":
        cells = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
        d = S.terrain(n, cells=cells, drop=0.3)
    elif kind == "bin":
        d = S.mixed_bin(n)
    else:
        d = S.pyramid(n)
    print(f"scene {d.name} built in {time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    ctx = Context(d, max_pairs=8 * d.n + 4096, max_manifolds=6 * d.n + 4096)
    print(f"upload {time.time() - t0:.1f}s", flush=True)
    for k in range(steps):
        t1 = time.time()
        ctx.step()
        ctx.sync()
        wall = (time.time() - t1) * 1e3
        if k % 10 == 0 or k == steps - 1:
            t = ctx.timings(); c = ctx.counts()
            print(f"step {k}: wall {wall:.2f} ms | dev total {t.total:.2f} bp {t.broadphase:.2f} np {t.narrowphase:.2f} build {t.contact_build:.2f} solve {t.solve:.2f}"
                  f" | pairs {c.n_pairs} manifolds {c.n_manifolds} points {c.n_points} colors {c.n_colors} overflow {c.n_overflow}", flush=True)
    p, q, v, w = ctx.get_state()
    print("min y", p[:, 1].min(), "max |v|", np.abs(v).max(), "nan", np.isnan(p).any())

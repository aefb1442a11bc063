"""Per-colour timing of the contact passes of k_substep_solve on the 1M-body terrain scene (or a smaller one): size of each
colour, microseconds per phase, and the algorithmic GB/s inside that phase."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from physecs_b200 import scenes as S
from physecs_b200.capi import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 150
d = S.terrain(n) if n == 1_000_000 else S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.05)), drop=0.3)
ctx = Context(d, max_pairs=8 * d.n, max_manifolds=6 * d.n)
for _ in range(settle):
    ctx.step()
ctx.sync()
ctx.set_profile(True)
steps = 20
for _ in range(steps):
    ctx.step()
ctx.sync()
m = ctx.manifolds()
hist = np.bincount(m["color"], minlength=64)
pts = np.bincount(m["color"], weights=m["num_points"], minlength=64)
prof = ctx.profile_colors()
tot = 0.0
for c in range(64):
    ms, cnt = prof[c]
    if not cnt:
        continue
    us = 1e3 * ms / cnt
    bytes_ = hist[c] * 180 + pts[c] * 140          # contact solve pass
    real = hist[c] * (16 + 16 + 8 + 8 + 64 + 64) + pts[c] * 112   # header, normal, L r/w, two velocity records r/w, seven row vectors
    tot += ms / steps
    print(f"colour {c:2d}: {hist[c]:8d} manifolds {int(pts[c]):8d} points  {us:8.2f} us/phase  {cnt / steps:5.1f} phases/step  algorithmic {bytes_ / us / 1e3:7.0f} GB/s  moved ~{real / us / 1e3:7.0f} GB/s")
print(f"contact passes {tot:.3f} ms/step;", {k: (round(v[0] / steps, 4), v[1] / steps) for k, v in ctx.profile().items()})

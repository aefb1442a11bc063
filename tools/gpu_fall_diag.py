"""Do the bodies of a walled bin land?  min / mean height and counts over time (tools/gpu_fall_diag.py 20000)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
n = int(sys.argv[1])
d = S.mixed_bin(n)
ctx = Context(d, max_pairs=64 * d.n + 4096, max_manifolds=16 * d.n + 4096)
for k in range(121):
    ctx.step()
    if k % 20 == 0:
        P = ctx.get_state()[0]; c = ctx.counts()
        print(f"step {k}: min y {P[:, 1].min():.2f} mean y {P[:, 1].mean():.2f} pairs {c.n_pairs} manifolds {c.n_manifolds} colours {c.n_colors} islands {ctx.island_stats()} bins {ctx.bin_counts() if hasattr(ctx, 'bin_counts') else ''}", flush=True)
ctx.close()

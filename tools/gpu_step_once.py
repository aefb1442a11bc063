"""Run a scene for `settle` steps, then a few profiled steps (cudaProfilerStart/Stop range) -- for ncu launch lists of small scenes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physecs_b200 import scenes as S
from physecs_b200.capi import Context, load_library
name = sys.argv[1]; n = int(sys.argv[2]); settle = int(sys.argv[3]); steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mk = {"ragdolls": lambda: S.ragdolls(n), "mixed": lambda: S.mixed_bin(n), "convex": lambda: S.convex_pile(n), "pyramid": lambda: S.pyramid(n)}[name]
d = mk()
ctx = Context(d, max_pairs=64 * d.n, max_manifolds=16 * d.n)
for _ in range(settle):
    ctx.step()
ctx.sync()
lib = load_library()
lib.pb_profiler_range(1)
for _ in range(steps):
    ctx.step()
ctx.sync()
lib.pb_profiler_range(0)
print(ctx.counts().n_manifolds, ctx.island_stats())

"""Plain step timing of a scene (no in-kernel profiling): ms/step over a timed loop and the per-stage CUDA-event split of the last steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
name = sys.argv[1]; n = int(sys.argv[2]); settle = int(sys.argv[3]); modes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [2]
mk = {"ragdolls": lambda: S.ragdolls(n), "mixed": lambda: S.mixed_bin(n), "convex": lambda: S.convex_pile(n), "pyramid": lambda: S.pyramid(n),
      "terrain": lambda: S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.05)), drop=0.3)}[name]
d = mk()
for mode in modes:
    ctx = Context(d, max_pairs=64 * d.n, max_manifolds=16 * d.n)
    ctx.set_islands(mode)
    for _ in range(settle):
        ctx.step()
    ctx.sync()
    steps = 100
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.step()
    ctx.sync()
    wall = (time.perf_counter() - t0) / steps * 1e3
    acc = np.zeros(6)
    for _ in range(10):
        ctx.step()
        t = ctx.timings()
        acc += np.array([t.broadphase, t.narrowphase, t.contact_build, t.solve, t.total, t.solve_kernel])
    acc /= 10
    c = ctx.counts()
    print(f"{d.name} islands={mode} {ctx.island_stats()}: {wall:.3f} ms/step; broad {acc[0]:.3f} narrow {acc[1]:.3f} build {acc[2]:.3f} solve {acc[3]:.3f} total {acc[4]:.3f}; manifolds {c.n_manifolds} colors {c.n_colors} launches/step {ctx.launches() / (settle + steps + 10):.1f}")
    ctx.close()

"""Phase breakdown of a scene on the GPU: pb_timings per stage + in-kernel phase stamps of k_substep_solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from physecs_b200 import scenes as S
from physecs_b200.capi import Context

name = sys.argv[1] if len(sys.argv) > 1 else "ragdolls"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mk = {"ragdolls": lambda: S.ragdolls(n), "mixed": lambda: S.mixed_bin(n), "convex": lambda: S.convex_pile(n), "pyramid": lambda: S.pyramid(n),
      "terrain": lambda: S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.05)), drop=0.3)}[name]
d = mk()
ctx = Context(d, max_pairs=64 * d.n, max_manifolds=16 * d.n)
settle = int(sys.argv[3]) if len(sys.argv) > 3 else 120
for _ in range(settle):
    ctx.step()
ctx.sync()
ctx.set_profile(True)
steps = 50
acc = np.zeros(6)
import time
t0 = time.perf_counter()
for _ in range(steps):
    ctx.step()
ctx.sync()
wall = (time.perf_counter() - t0) / steps * 1e3
for _ in range(10):
    ctx.step()
    t = ctx.timings()
    acc += np.array([t.broadphase, t.narrowphase, t.contact_build, t.solve, t.total, t.solve_kernel])
acc /= 10
prof = ctx.profile()
c = ctx.counts()
print(f"{d.name}: bodies={ctx.n_dyn} pairs={c.n_pairs} manifolds={c.n_manifolds} points={c.n_points} colors={c.n_colors} joints={len(d.joints)}")
print(f"wall ms/step (async loop) {wall:.3f}; stage ms: broad {acc[0]:.3f} narrow {acc[1]:.3f} build {acc[2]:.3f} solve {acc[3]:.3f} total {acc[4]:.3f} (solve kernels {acc[5]:.3f})")
for k, (ms, cnt) in prof.items():
    print(f"  {k:14s} {ms / steps:8.4f} ms/step  phases/step {cnt / steps:6.1f}  us/phase {1e3 * ms / max(cnt, 1):7.2f}")
pc = ctx.profile_colors()
for k, nm in enumerate(["local NGS", "local contact", "local joint"]):
    ms, cnt = pc[61 + k]
    print(f"  CTA0 {nm:14s} {ms / steps:8.4f} ms/step  phases/step {cnt / steps:6.1f}  us/phase {1e3 * ms / max(cnt, 1):7.2f}")
print("launches/step", ctx.launches() / (settle + steps + 10))

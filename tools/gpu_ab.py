"""A/B timing of runtime knobs on one scene: tools/gpu_ab.py terrain 1000000 150 PB_MESH_LIGHT=0 PB_MESH_LIGHT=1 PB_MESH_LIGHT=2
Each variant is a comma-separated list of ENV=value settings read at context creation; the scene description is built once."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
name = sys.argv[1]; n = int(sys.argv[2]); settle = int(sys.argv[3]); variants = sys.argv[4:] or [""]
mk = {"ragdolls": lambda: S.ragdolls(n), "mixed": lambda: S.mixed_bin(n), "convex": lambda: S.convex_pile(n), "pyramid": lambda: S.pyramid(n),
      "terrain": lambda: S.terrain(n, drop=0.3) if n == 1_000_000 else S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.05)), drop=0.3)}[name]
d = mk()
for var in variants:
    kv = [x.split("=") for x in var.split(",") if x]
    for k, v in kv:
        os.environ[k] = v
    ctx = Context(d, max_pairs=(8 if name == 'terrain' else 64) * d.n + 4096, max_manifolds=(6 if name == 'terrain' else 16) * d.n + 4096)
    for _ in range(settle):
        ctx.step()
    ctx.sync()
    best = 1e9
    for rep in range(3):
        steps = 50
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.step()
        ctx.sync()
        best = min(best, (time.perf_counter() - t0) / steps * 1e3)
    acc = np.zeros(6)
    for _ in range(10):
        ctx.step()
        t = ctx.timings()
        acc += np.array([t.broadphase, t.narrowphase, t.contact_build, t.solve, t.total, t.solve_kernel])
    acc /= 10
    c = ctx.counts()
    pos = ctx.get_state()[0] if hasattr(ctx, "get_state") else None
    chk = float(np.abs(pos).sum()) if pos is not None else 0.0
    print(f"{d.name} [{var or 'default'}]: best {best:.3f} ms/step; broad {acc[0]:.3f} narrow {acc[1]:.3f} build {acc[2]:.3f} solve {acc[3]:.3f} total {acc[4]:.3f}; "
          f"pairs {c.n_pairs} manifolds {c.n_manifolds} colours {c.n_colors} islands {ctx.island_stats()} checksum {chk:.6e}", flush=True)
    ctx.close()
    for k, _ in kv:
        os.environ.pop(k, None)

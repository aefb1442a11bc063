"""Host time spent inside pb_step (enqueue only) next to the device time of the step, per variant: tells when a small scene is bound by
the host's launch rate rather than by the device.  usage: python tools/gpu_cpu_enqueue.py <scene> <n> <settle> VAR=..,VAR=.. ..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
name = sys.argv[1]; n = int(sys.argv[2]); settle = int(sys.argv[3]); variants = sys.argv[4:] or [""]
mk = {"ragdolls": lambda: S.ragdolls(n), "mixed": lambda: S.mixed_bin(n), "pyramid": lambda: S.pyramid(n)}[name]
d = mk()
for var in variants:
    kv = [x.split("=") for x in var.split(",") if x]
    for k, v in kv:
        os.environ[k] = v
    ctx = Context(d, max_pairs=64 * d.n + 4096, max_manifolds=16 * d.n + 4096)
    for _ in range(settle):
        ctx.step()
    ctx.sync()
    l0 = ctx.launches()
    cpu = []
    t0 = time.perf_counter()
    for _ in range(200):
        a = time.perf_counter(); ctx.step(); cpu.append(time.perf_counter() - a)
    ctx.sync()
    wall = (time.perf_counter() - t0) / 200 * 1e3
    t = ctx.timings()
    print(f"{d.name} [{var or 'default'}]: wall {wall:.3f} ms/step, host time in pb_step median {np.median(cpu) * 1e3:.3f} ms (p90 {np.percentile(cpu, 90) * 1e3:.3f}), "
          f"device total of the last step {t.total:.3f} ms, {(ctx.launches() - l0) / 200:.1f} launches/step, islands {ctx.island_stats()}", flush=True)
    ctx.close()
    for k, _ in kv:
        os.environ.pop(k, None)

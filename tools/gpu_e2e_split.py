"""Where the end-to-end step time goes at 1M bodies: pb_step alone, pb_set_state + pb_step, pb_get_state alone, all three."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = S.terrain(n, drop=0.3) if n == 1_000_000 else S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.05)), drop=0.3)
ctx = Context(d, max_pairs=8 * d.n + 4096, max_manifolds=6 * d.n + 4096)
for _ in range(150):
    ctx.step()
ctx.sync()
nd = ctx.n_dyn
bufs = [torch.empty((nd, k), dtype=torch.float32).pin_memory().numpy() for k in (3, 4, 3, 3)]
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
lib = ctx.lib
lib.pb_get_state(ctx.ctx, *[fp(b) for b in bufs])
def timeit(fn, k=30):
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / k * 1e3
step = lambda: lib.pb_step(ctx.ctx, C.c_float(d.dt), d.substeps, d.iterations, C.c_float(d.gravity))
setst = lambda: lib.pb_set_state(ctx.ctx, nd, *[fp(b) for b in bufs])
getst = lambda: lib.pb_get_state(ctx.ctx, *[fp(b) for b in bufs])
print("pb_step                      %.3f ms" % timeit(step))
print("pb_set_state (+sync)         %.3f ms" % timeit(lambda: (setst(), ctx.sync())))
print("pb_get_state                 %.3f ms" % timeit(getst))
print("set_state + step             %.3f ms" % timeit(lambda: (setst(), step())))
print("step + get_state             %.3f ms" % timeit(lambda: (step(), getst())))
print("set_state + step + get_state %.3f ms" % timeit(lambda: (setst(), step(), getst())))

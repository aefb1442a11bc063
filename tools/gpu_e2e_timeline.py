"""Host-side timeline of one iteration of the chunked round-trip loop (bench.py e2e): when does every call return?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = S.terrain(n, drop=0.3) if n == 1_000_000 else S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.05)), drop=0.3)
ctx = Context(d, max_pairs=8 * d.n + 4096, max_manifolds=6 * d.n + 4096)
for _ in range(150):
    ctx.step()
ctx.sync()
N = ctx.n_dyn
pin = lambda w: torch.empty((N, w), dtype=torch.float32).pin_memory()
tp, tq, tv, tw = pin(3), pin(4), pin(3), pin(3)
pos, quat, vel, ang = tp.numpy(), tq.numpy(), tv.numpy(), tw.numpy()
fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
lib, h = ctx.lib, ctx.ctx
lib.pb_get_state(h, fp(pos), fp(quat), fp(vel), fp(ang))
first, count = C.c_int(), C.c_int()
NULLF = C.POINTER(C.c_float)()
lib.pb_set_readback_order(h, 1)
lib.pb_step_begin(h)
logs = []
T0 = time.perf_counter()
for k in range(30):
    ev = []; t = lambda name: ev.append((name, (time.perf_counter() - T0) * 1e3))
    t("start")
    for half in (0, 1):
        for c in range(CH if k > 0 else 1):
            if k > 0:
                (lib.pb_get_state_wait if half else lib.pb_get_state_wait_poses)(h, c, C.byref(first), C.byref(count)); f0, cnt = first.value, count.value
                t(f"{'vel' if half else 'pose'}{c}")
            else:
                f0, cnt = 0, N
            off = lambda x, w: C.cast(C.c_void_p(x.ctypes.data + 4 * w * f0), C.POINTER(C.c_float))
            if half == 0:
                lib.pb_set_state_rows(h, f0, cnt, off(pos, 3), off(quat, 4), NULLF, NULLF)
            else:
                lib.pb_set_state_rows(h, f0, cnt, NULLF, NULLF, off(vel, 3), off(ang, 3))
        if half == 0:
            lib.pb_step_narrowphase(h); t("narrowphase enqueued")
    t("uploads enqueued")
    lib.pb_step(h, C.c_float(d.dt), d.substeps, d.iterations, C.c_float(d.gravity)); t("pb_step returned")
    lib.pb_get_state_begin(h, fp(pos), fp(quat), fp(vel), fp(ang), CH); t("get_state_begin returned")
    lib.pb_step_begin(h); t("step_begin returned")
    logs.append(ev)
    if k in (20, 21):
        for c in range(CH):
            lib.pb_get_state_wait(h, c, C.byref(first), C.byref(count))
        tm = ctx.timings()
        print(f"   device stages of step {k}: broad {tm.broadphase:.3f} narrow {tm.narrowphase:.3f} build {tm.contact_build:.3f} solve {tm.solve:.3f} total {tm.total:.3f}")
for c in range(CH):
    lib.pb_get_state_wait(h, c, C.byref(first), C.byref(count))
total = (time.perf_counter() - T0) * 1e3
print(f"{total / 30:.3f} ms per iteration")
for ev in logs[20:22]:
    base = ev[0][1]
    print("  ".join(f"{name} +{ts - base:.2f}" for name, ts in ev))
ctx.close()

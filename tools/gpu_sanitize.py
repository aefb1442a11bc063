"""Small scenes through every device path, for compute-sanitizer (tools/gpu_sanitize.sh): island sweeps on / off, the whole-step kernel
(grid-barrier and group-by-group forms), the cooperative sort, tree and all-pairs broadphase, the spill
kernels, deterministic colouring.  No oracle: the sanitizer is the checker here."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physecs_b200 import scenes as S
from physecs_b200.capi import Context

def run(name, desc, steps, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        ctx = Context(desc)
        for _ in range(steps):
            ctx.step()
        c = ctx.counts()
        print(f"{name}: manifolds {c.n_manifolds} colours {c.n_colors} cause {c.cause:#x} islands {ctx.island_stats()}", flush=True)
        ctx.close()
    finally:
        for k in env:
            os.environ.pop(k, None)

run("pyramid (one launch per step, grid barriers)", S.pyramid(60), 4)
run("ragdolls, group by group", S.ragdolls(6), 45, PB_ISLANDS=1)
run("ragdolls, per-substep launches + island sweeps", S.ragdolls(6), 45, PB_ISLANDS=1, PB_FUSED=0)
run("mixed bin, tree broadphase (packet walk) + cooperative sort + device-wide sweeps", S.mixed_bin(9000, spacing=0.8), 6, PB_BRUTE_FORCE_MAX=0, PB_ISLANDS=0, PB_FUSED=0)
run("terrain, tree broadphase + big-static list", S.terrain(600, cells=32, drop=0.05), 12, PB_BRUTE_FORCE_MAX=0)
run("terrain mixed (all shape kinds on a mesh)", S.terrain_mixed(300, cells=24, drop=0.05), 10)
run("big shapes on a fine mesh (mesh spill kernel)", S.big_on_fine_mesh(cells=40), 3)
run("degenerate convex pairs (GJK / EPA spill kernel)", S.degenerate_convex(), 3)
run("convex pile, deterministic colouring", S.convex_pile(300, mix_prims=True), 30, PB_DETERMINISTIC=1)
run("joint zoo + overflow bucket", S.joint_star(12), 10)
run("800 ragdoll scenes: tree + tile probe on the first step, all-pairs kernel with tile culling from the second", S.ragdolls(800), 4)
run("pyramid as one thread-block cluster (hardware cluster barrier)", S.pyramid(60), 4, PB_ISLANDS=0, PB_FUSED=1)
run("mixed bin, small and big islands in one step (per-CTA sweeps + device-wide colours)", S.mixed_bin(2500, spacing=0.8), 30, PB_ISLAND_LOCAL_MAX=40, PB_FUSED=0, PB_ISLANDS=1)


def round_trip():
    # read-back on the read stream beside the next step's broadphase / narrowphase, uploads on the copy stream (bench.py e2e loop)
    import ctypes as C
    import numpy as np
    d = S.mixed_bin(1500, spacing=0.8)
    ctx = Context(d)
    n = ctx.n_dyn
    pos, quat, vel, ang = (np.zeros((n, w), np.float32) for w in (3, 4, 3, 3))
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    lib, h = ctx.lib, ctx.ctx
    NULLF = C.POINTER(C.c_float)()
    first, count = C.c_int(), C.c_int()
    lib.pb_get_state(h, fp(pos), fp(quat), fp(vel), fp(ang)); lib.pb_set_readback_order(h, 1); lib.pb_step_begin(h)
    for k in range(12):
        if k:
            lib.pb_get_state_wait_poses(h, 0, C.byref(first), C.byref(count))
        lib.pb_set_state_rows(h, 0, n, fp(pos), fp(quat), NULLF, NULLF); lib.pb_step_narrowphase(h)
        if k:
            lib.pb_get_state_wait(h, 0, C.byref(first), C.byref(count))
        lib.pb_set_state_rows(h, 0, n, NULLF, NULLF, fp(vel), fp(ang))
        assert lib.pb_step(h, C.c_float(d.dt), d.substeps, d.iterations, C.c_float(d.gravity)) == 0
        assert lib.pb_get_state_begin(h, fp(pos), fp(quat), fp(vel), fp(ang), 1) == 0
        lib.pb_step_begin(h)
    lib.pb_get_state_wait(h, 0, C.byref(first), C.byref(count))
    print(f"round trip beside the step: manifolds {ctx.counts().n_manifolds} rows read {count.value}", flush=True)
    ctx.close()


round_trip()
print("sanitize scenes done")

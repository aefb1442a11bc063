"""-m gpu: deterministic mode (PB_DETERMINISTIC / pb_set_deterministic).  The reference with numThreads = 0 is reproducible
(src/ThreadPool.cpp:30-46); the device's default colouring is a race, so every run takes another valid colour order.  With fixed
colour priorities two runs of the same scene must give bit-identical states, and the mode must pass the same gates as the default."""
import numpy as np
import pytest

from physecs_b200 import scenes as S
from tests import parity
from tests.test_gpu_fullsize import check_colouring

pytestmark = pytest.mark.gpu


def _run(desc, steps, islands):
    from physecs_b200.capi import Context
    ctx = Context(desc, max_pairs=32 * desc.n + 4096, max_manifolds=12 * desc.n + 4096)
    try:
        ctx.set_deterministic(True)
        ctx.set_islands(islands)
        for _ in range(steps):
            ctx.step()
        check_colouring(ctx, desc)
        return ctx.get_state(), ctx.counts()
    finally:
        ctx.close()


@pytest.mark.parametrize("name,maker,steps", [
    ("C3_convex_pile_250k", lambda: S.convex_pile(250_000), 200),
    ("mixed_bin_20k", lambda: S.mixed_bin(20_000), 200),
    ("ragdolls_256", lambda: S.ragdolls(256), 150),
    ("terrain_mixed_3k", lambda: S.terrain_mixed(3000, cells=64), 150),
], ids=lambda x: x if isinstance(x, str) else None)
def test_two_runs_are_bit_identical(name, maker, steps):
    d = maker()
    (a, ca), (b, cb) = _run(d, steps, 2), _run(d, steps, 2)
    assert ca.n_manifolds == cb.n_manifolds and ca.n_colors == cb.n_colors and ca.n_manifolds > 0
    for x, y, what in zip(a, b, ("pos", "quat", "vel", "angvel")):
        assert np.array_equal(x.view(np.int32), y.view(np.int32)), f"{name}: {what} differs between two deterministic runs"


def test_islands_on_and_off_agree_in_deterministic_mode():
    """with fixed colours the island grouping only changes which CTA sweeps a manifold, never the order on a body"""
    d = S.mixed_bin(6000, spacing=0.8)
    (a, _), (b, _) = _run(d, 80, 0), _run(d, 80, 1)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.int32), y.view(np.int32))


@pytest.mark.parametrize("maker", [lambda: S.pyramid(120), lambda: S.mixed_bin(1500, spacing=0.8), lambda: S.ragdolls(6), lambda: S.convex_pile(400, mix_prims=True)])
def test_gates_in_deterministic_mode(maker, monkeypatch):
    monkeypatch.setenv("PB_DETERMINISTIC", "1")
    s = parity.run_gates(maker(), steps=8)
    assert s["steps"] == 8 and s["worst_manifold"] <= parity.TOL


def test_chunked_round_trip_beside_the_step_equals_plain_steps():
    """The host-authoritative loop with the transfers beside the device's work -- pb_get_state_begin (chunks, read stream), pb_step_begin
    (the next broadphase at once), every chunk uploaded again as it arrives (pb_set_state_rows), pb_step -- against plain pb_step calls
    on a second context: bit-identical states (deterministic colours), i.e. the overlap changes nothing the step sees."""
    import ctypes as C
    import torch
    from physecs_b200.capi import Context
    d = S.mixed_bin(20_000)
    a = Context(d, max_pairs=32 * d.n + 4096, max_manifolds=12 * d.n + 4096)
    b = Context(d, max_pairs=32 * d.n + 4096, max_manifolds=12 * d.n + 4096)
    try:
        for c in (a, b):
            c.set_deterministic(True)
        n = a.n_dyn
        pin = lambda w: torch.empty((n, w), dtype=torch.float32).pin_memory()
        tp, tq, tv, tw = pin(3), pin(4), pin(3), pin(3)
        pos, quat, vel, ang = tp.numpy(), tq.numpy(), tv.numpy(), tw.numpy()
        fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        lib, h = b.lib, b.ctx
        assert lib.pb_get_state(h, fp(pos), fp(quat), fp(vel), fp(ang)) == 0
        CH, steps = 5, 70
        first, count = C.c_int(), C.c_int()
        NULLF = C.POINTER(C.c_float)()
        assert lib.pb_set_readback_order(h, 1) == 0
        assert lib.pb_step_begin(h) == 0
        for k in range(steps):
            a.step()
            for half in (0, 1):
                for c in range(CH if k > 0 else 1):
                    if k > 0:
                        assert (lib.pb_get_state_wait if half else lib.pb_get_state_wait_poses)(h, c, C.byref(first), C.byref(count)) == 0
                        f0, cnt = first.value, count.value
                    else:
                        f0, cnt = 0, n
                    if cnt:
                        off = lambda x, w: C.cast(C.c_void_p(x.ctypes.data + 4 * w * f0), C.POINTER(C.c_float))
                        if half == 0:
                            assert lib.pb_set_state_rows(h, f0, cnt, off(pos, 3), off(quat, 4), NULLF, NULLF) == 0
                        else:
                            assert lib.pb_set_state_rows(h, f0, cnt, NULLF, NULLF, off(vel, 3), off(ang, 3)) == 0
                if half == 0:
                    assert lib.pb_step_narrowphase(h) == 0
            assert lib.pb_step(h, C.c_float(d.dt), d.substeps, d.iterations, C.c_float(d.gravity)) == 0, lib.pb_last_error(h)
            assert lib.pb_get_state_begin(h, fp(pos), fp(quat), fp(vel), fp(ang), CH) == 0
            if k + 1 < steps:
                assert lib.pb_step_begin(h) == 0
        seen = 0
        for c in range(CH):
            assert lib.pb_get_state_wait(h, c, C.byref(first), C.byref(count)) == 0
            seen += count.value
        assert seen == n
        A = a.get_state()
        for x, y, what in zip(A, (pos, quat, vel, ang), ("pos", "quat", "vel", "angvel")):
            assert np.array_equal(x.view(np.int32), y.view(np.int32)), what
        B = b.get_state()            # ... and the device holds what was read
        for x, y in zip(A, B):
            assert np.array_equal(x.view(np.int32), y.view(np.int32))
        assert a.counts().n_manifolds > 10000
    finally:
        a.close(); b.close()


def test_scene_edit_between_narrowphase_and_step_restarts_the_step():
    """pb_step_narrowphase enqueues half a step; a scene edit behind it (rows moved: bounds change) makes pb_step start over.  Same result
    as the edit followed by a plain pb_step, over several steps (the abandoned narrowphase leaves nothing behind)."""
    from physecs_b200.capi import Context
    d = S.mixed_bin(3000, spacing=0.8)
    a, b = Context(d), Context(d)
    try:
        dyn = d.dynamic_entities()
        for c in (a, b):
            c.set_deterministic(True)
            for _ in range(40):
                c.step()
        for k in range(6):
            P, Q, _, _ = a.get_state_entities()
            ents = dyn[k::7][:50]
            newP = P[ents] + np.float32([0.0, 0.5, 0.0]); newQ = Q[ents]
            a.move_rows(ents, newP, newQ)
            a.step()
            assert b.lib.pb_step_narrowphase(b.ctx) == 0
            b.move_rows(ents, newP, newQ)
            b.step()
            b.step(); a.step()
            for x, y, what in zip(a.get_state(), b.get_state(), ("pos", "quat", "vel", "angvel")):
                assert np.array_equal(x.view(np.int32), y.view(np.int32)), (k, what)
        assert a.counts().n_manifolds == b.counts().n_manifolds and a.counts().n_manifolds > 1000
    finally:
        a.close(); b.close()

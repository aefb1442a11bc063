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

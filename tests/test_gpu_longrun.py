"""-m gpu: long-horizon statistics (north_star: "long-horizon runs are additionally checked statistically on energy drift and
maximum penetration").  Device and oracle both free-run from the same initial state with their OWN constraint orders
(the device's colour order is non-deterministic), so states diverge chaotically and only distributions are compared.
The scenes are walled bins: open piles (pyramid, convex pile on a finite floor) keep collapsing / spilling for seconds in the
reference too, which leaves nothing stationary to compare."""
import numpy as np
import pytest

from physecs_b200 import scenes as S
from physecs_b200.capi import Context

pytestmark = pytest.mark.gpu


def _energy(d, pos, quat, vel, ang):
    dyn = d.dynamic_entities()
    m = 1.0 / d.inv_mass[dyn].astype(np.float64)
    v, w, q = vel[dyn].astype(np.float64), ang[dyn].astype(np.float64), quat[dyn].astype(np.float64)
    R = S.quat_to_mat(q) if q.ndim == 1 else np.stack([S.quat_to_mat(x) for x in q])
    Iinv = d.inv_inertia[dyn].astype(np.float64).reshape(-1, 3, 3).transpose(0, 2, 1)      # column-major -> row-major
    wl = np.einsum("nji,nj->ni", R, w)                                                       # body-frame angular velocity
    I = np.linalg.inv(Iinv)
    rot = 0.5 * np.einsum("ni,nij,nj->n", wl, I, wl)
    com_y = pos[dyn, 1].astype(np.float64) + np.einsum("nij,nj->ni", R, d.com[dyn].astype(np.float64))[:, 1]
    return float(np.sum(0.5 * m * np.sum(v * v, 1) + rot + m * d.gravity * com_y))


def _penetration(m):
    """depth of every contact point: positive when the witness points have crossed along the normal."""
    if len(m["keys"]) == 0:
        return np.zeros(0)
    n = m["normal"][:, None, :]
    dep = np.sum((m["points"][:, :, 0, :] - m["points"][:, :, 1, :]) * n, axis=2)
    mask = np.arange(4)[None, :] < m["num_points"][:, None]
    return dep[mask]


@pytest.mark.parametrize("maker,steps", [(lambda: S.mixed_bin(1500, spacing=0.8), 360), (lambda: S.mixed_bin(600, spacing=0.8, seed=0x51), 300)])
def test_energy_and_penetration_statistics(maker, steps):
    from oracle.ref import RefScene
    d = maker()
    ref = RefScene(d, 0, hashfix=True)
    ctx = Context(d)
    e_dev, e_ref = [], []
    sample = range(0, steps, 20)
    for k in range(steps):
        ctx.step(); ref.simulate()
        if k in sample or k == steps - 1:
            P, Q, V, W = ctx.get_state_entities()
            e_dev.append(_energy(d, P, Q, V, W))
            e_ref.append(_energy(d, *ref.get_state()))
    e_dev, e_ref = np.array(e_dev), np.array(e_ref)
    drop = abs(e_ref[0] - e_ref[-1]) + 1e-9
    # energy is dissipated along the same curve: compare at every sample, relative to the total energy the scene sheds
    # (a stack that merely settles sheds little, so the bound is also expressed relative to the total energy)
    bound = max(0.05 * drop, 3e-3 * abs(e_ref[0]))
    assert np.max(np.abs(e_dev - e_ref)) < bound, (e_dev, e_ref)
    # no energy is created once the scene rests (last third of the run): same drift as the reference shows
    tail, tail_ref = e_dev[2 * len(e_dev) // 3:], e_ref[2 * len(e_ref) // 3:]
    assert tail.max() - tail.min() < max(0.02 * drop, 2.0 * (tail_ref.max() - tail_ref.min()) + 1e-3 * abs(e_ref[0]))
    # penetration: same distribution of contact depths at the end of the run
    gm, rm = ctx.manifolds(), ref.narrowphase(ref.pairs())
    pd, pr = _penetration(gm), _penetration(rm)
    # same contact graph size; the number of POINTS per resting face contact flickers with the 1e-4 depth gate, so it is looser
    assert abs(len(gm["keys"]) - len(rm["keys"])) < 0.1 * len(rm["keys"]) + 10
    assert len(pd) > 50 and abs(len(pd) - len(pr)) < 0.35 * len(pr) + 10
    assert pd.max() < 0.06 and abs(pd.max() - pr.max()) < 0.02, (pd.max(), pr.max())
    assert abs(pd.mean() - pr.mean()) < 0.003, (pd.mean(), pr.mean())
    # the bodies came to rest in the same place on average
    P, _, V, _ = ctx.get_state_entities()
    p, _, v, _ = ref.get_state()
    dyn = d.dynamic_entities()
    # (median: a few bodies that roll off the floor keep falling and would dominate a mean)
    assert abs(np.median(P[dyn, 1]) - np.median(p[dyn, 1])) < 0.03
    # ... and are equally calm: a collapsing stack always has a few movers (which ones is chaotic), so compare speed quantiles
    sd, sr = np.linalg.norm(V[dyn], axis=1), np.linalg.norm(v[dyn], axis=1)
    for qt in (50, 90):
        assert abs(np.percentile(sd, qt) - np.percentile(sr, qt)) < 0.15, (qt, np.percentile(sd, qt), np.percentile(sr, qt))
    assert np.percentile(sd, 90) < 0.5
    ctx.close(); ref.close()

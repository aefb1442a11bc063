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


@pytest.mark.parametrize("name,maker,steps", [
    ("ragdolls", lambda: S.ragdolls(24), 420),                                   # joints + contacts: bodies fall, fold up and come to rest
    ("terrain", lambda: S.terrain(1200, cells=48, drop=0.3), 420),               # spheres / capsules rolling on a triangle mesh
    ("terrain_mixed", lambda: S.terrain_mixed(600, cells=40), 360),              # all four shape kinds on the mesh
])
def test_long_horizon_joints_and_terrain(name, maker, steps):
    """The same statistical comparison on the scene families the bin scenes do not cover: jointed bodies in contact (ragdolls) and
    bodies on a triangle mesh.  Device and oracle free-run with their own constraint orders; trajectories differ, distributions agree."""
    from oracle.ref import RefScene
    d = maker()
    ref = RefScene(d, 0, hashfix=True)
    ctx = Context(d)
    e_dev, e_ref = [], []
    for k in range(steps):
        ctx.step(); ref.simulate()
        if k % 20 == 0 or k == steps - 1:
            P, Q, V, W = ctx.get_state_entities()
            assert np.isfinite(P).all() and np.isfinite(V).all() and np.isfinite(Q).all() and np.isfinite(W).all()
            e_dev.append(_energy(d, P, Q, V, W))
            e_ref.append(_energy(d, *ref.get_state()))
    e_dev, e_ref = np.array(e_dev), np.array(e_ref)
    drop = abs(e_ref[0] - e_ref[-1]) + 1e-9
    if name == "ragdolls":
        # The reference's ragdolls do not come to rest: jointed limbs in contact with the ground keep gaining energy in the ORACLE too
        # (SURVEY.md 7.8 flags the configuration; here the total triples within 7 s on both sides).  What can be asked of the device
        # is that it misbehaves the same way: energy curves of the same size, never far above the oracle's.
        # The growth is chaotic (two runs of the oracle with different constraint orders differ as much), so: same order of magnitude
        # at every sample, and the device never runs away from the oracle.
        ratio = np.abs(e_dev) / np.maximum(np.abs(e_ref), 1e-9)
        assert ratio.min() > 1.0 / 3.0 and ratio.max() < 3.0, (e_dev, e_ref)
        assert e_dev.max() < 2.0 * e_ref.max() + abs(e_ref[0])
    else:
        # the energy shed over the run agrees (rolling bodies on a slope keep shedding: compare the curves loosely, the end points tighter)
        assert np.max(np.abs(e_dev - e_ref)) < max(0.12 * drop, 5e-3 * abs(e_ref[0])), (e_dev, e_ref)
        assert abs(e_dev[-1] - e_ref[-1]) < max(0.06 * drop, 3e-3 * abs(e_ref[0])), (e_dev[-1], e_ref[-1])
        # no energy is created: the device's curve never rises above its start by more than the oracle's does
        assert e_dev.max() - e_dev[0] < max(e_ref.max() - e_ref[0], 0.0) + 2e-3 * abs(e_ref[0]) + 0.01 * drop
    gm, rm = ctx.manifolds(), ref.narrowphase(ref.pairs())
    pd, pr = _penetration(gm), _penetration(rm)
    P, _, V, _ = ctx.get_state_entities()
    p, _, v, _ = ref.get_state()
    dyn = d.dynamic_entities()
    if name != "ragdolls":       # (thrashing ragdolls have no stationary contact set or rest pose to compare)
        assert abs(len(gm["keys"]) - len(rm["keys"])) < 0.15 * len(rm["keys"]) + 10
        assert pd.size and pr.size and pd.max() < 0.08 and abs(pd.max() - pr.max()) < 0.03, (pd.max(), pr.max())
        assert abs(pd.mean() - pr.mean()) < 0.004, (pd.mean(), pr.mean())
        assert abs(np.median(P[dyn, 1]) - np.median(p[dyn, 1])) < 0.05
        sd, sr = np.linalg.norm(V[dyn], axis=1), np.linalg.norm(v[dyn], axis=1)
        assert abs(np.percentile(sd, 50) - np.percentile(sr, 50)) < 0.2 and abs(np.percentile(sd, 90) - np.percentile(sr, 90)) < 0.4
    if name == "ragdolls":
        # joints hold: the distance between the bodies of every joint stays what the oracle's is (anchors coincide up to solver slack)
        def gaps(pos, quat):
            out = []
            for (t, e0, a0p, a0q, e1, a1p, a1q, prm) in d.joints:
                w0 = pos[e0] + S.qrot(quat[e0].astype(np.float64), a0p.astype(np.float64))
                w1 = pos[e1] + S.qrot(quat[e1].astype(np.float64), a1p.astype(np.float64))
                out.append(np.linalg.norm(w0 - w1))
            return np.array(out)
        _, Qd, _, _ = ctx.get_state_entities()
        pr_, qr_, _, _ = ref.get_state()
        gd, gr = gaps(P, Qd), gaps(pr_, qr_)
        # (quantiles, not the maximum: in this configuration single ragdolls blow up in the ORACLE as well -- which ones, and how far,
        # depends on the constraint order; seen: device maximum 1.5 against the oracle's 0.24 in one run, the reverse in others)
        for qt in (50, 90):
            assert np.percentile(gd, qt) < max(3.0 * np.percentile(gr, qt), 0.02), (qt, np.percentile(gd, qt), np.percentile(gr, qt))
        assert gd.max() < max(20.0 * gr.max(), 5.0), (gd.max(), gr.max())
    ctx.close(); ref.close()

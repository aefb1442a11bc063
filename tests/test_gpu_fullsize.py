"""-m gpu: the three gates at BASELINE.json's FULL sizes (configs[1..4]).

Each scene is first settled on the device alone (so the contact graph is populated: piles, resting contacts, multi-point
manifolds), the settled state becomes a fresh scene description, and then device and oracle are created from it and
compared step by step exactly like the small scenes of test_gpu_gates.py: collider bounds bit-exact, candidate pair sets
identical, every manifold (key incl. triangle index, point count, normal, witness points), and the one-step solve of the
reference CPU solver fed the device's (colour, slot) order -- all bodies, all velocities.  Plus the size-independent
property that makes a parallel colour equal to a sequential sub-sweep: no two manifolds of a colour share a dynamic body.

The oracle (single-threaded reference) needs ~1 s/step at 100k bodies and ~10-20 s/step at 1M, so step counts are small.
"""
import copy

import numpy as np
import pytest

from physecs_b200 import scenes as S
from tests import parity

pytestmark = pytest.mark.gpu


def settled(desc, steps, caps=None):
    """Run `steps` steps on the device and return a copy of the description holding the reached state."""
    from physecs_b200.capi import Context
    ctx = Context(desc, **(caps or {}))
    try:
        for _ in range(steps):
            ctx.step()
        P, Q, V, W = ctx.get_state_entities()
        counts = ctx.counts()
        check_colouring(ctx, desc)
    finally:
        ctx.close()
    d = copy.copy(desc)
    d.pos, d.quat, d.vel, d.angvel = P, Q, V, W
    return d, counts


def check_colouring(ctx, desc):
    """Within one colour no dynamic body appears twice (static / kinematic sides do not count: they are never written)."""
    m = ctx.manifolds()
    if not len(m["keys"]):
        return
    dyn = (desc.flags & S.F_DYNAMIC) != 0
    kin = (desc.flags & S.F_KINEMATIC) != 0
    live = dyn & ~kin
    e0, e1, col = m["keys"][:, 0], m["keys"][:, 2], m["color"].astype(np.int64)
    ent = np.concatenate([e0[live[e0]], e1[live[e1]]]).astype(np.int64)
    c = np.concatenate([col[live[e0]], col[live[e1]]])
    key = c * (desc.n + 1) + ent
    assert len(np.unique(key)) == len(key), "a colour holds two manifolds on the same dynamic body"


CASES = [
    # (name, maker, settle steps, gated steps, minimum manifolds expected)
    ("C2_mixed_bin_100k", lambda: S.mixed_bin(100_000), 150, 3, 100_000),
    ("C3_convex_pile_250k", lambda: S.convex_pile(250_000), 90, 2, 50_000),
    ("C4_terrain_1M", lambda: S.terrain(1_000_000), 120, 2, 1_000_000),
    ("C5_ragdolls_4096", lambda: S.ragdolls(4096), 60, 3, 10_000),
]


@pytest.mark.parametrize("name,maker,settle,steps,min_manifolds", CASES, ids=[c[0] for c in CASES])
def test_full_size_gates(name, maker, settle, steps, min_manifolds):
    desc, counts = settled(maker(), settle)
    s = parity.run_gates(desc, steps=steps, bulk=True)
    assert s["steps"] == steps
    assert s["manifolds"] >= min_manifolds, f"{name}: only {s['manifolds']} manifolds after settling: the test checks little"
    assert s["worst_manifold"] <= parity.TOL
    print(name, s)

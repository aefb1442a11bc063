"""-m gpu: the three north_star gates through the C ABI, against the oracle, on seeded scenes."""
import numpy as np
import pytest

from physecs_b200 import scenes as S
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("maker,steps,need_contacts", [
    (lambda: S.pyramid(120), 40, True),
    (lambda: S.mixed_bin(1200, spacing=0.8), 60, True),
    (lambda: S.terrain(1500, cells=48, drop=0.3), 60, True),
    (lambda: S.joint_zoo(), 50, False),        # every joint type except gear; servo uses acos -> 1e-4 gate, the rest are bit-exact
    (lambda: S.ragdolls(8), 90, True),
    (lambda: S.convex_pile(300, mix_prims=True), 90, True),   # config 3 in miniature: GJK/EPA, all X-convex routines
    (lambda: S.terrain_mixed(800, cells=40), 90, True),        # sphere / capsule / box / convex vs triangle mesh        # config 5 in miniature: joints + contacts
])
def test_three_gates(maker, steps, need_contacts):
    desc = maker()
    s = parity.run_gates(desc, steps=steps)
    assert s["steps"] == steps
    assert s["manifolds"] > 0 or not need_contacts, "scene produced no contacts: the test checks nothing"
    assert s["worst_manifold"] <= parity.TOL

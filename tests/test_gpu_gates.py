"""-m gpu: the three north_star gates through the C ABI, against the oracle, on seeded scenes."""
import numpy as np
import pytest

from physecs_b200 import scenes as S
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("maker,steps", [
    (lambda: S.pyramid(120), 40),
    (lambda: S.mixed_bin(1200, spacing=0.8), 60),
    (lambda: S.terrain(1500, cells=48, drop=0.3), 60),
    (lambda: S.joint_zoo(), 50),          # every joint type except gear; servo uses acos -> 1e-4 gate, the rest are bit-exact
    (lambda: S.ragdolls(8), 90),          # config 5 in miniature: joints + contacts
])
def test_three_gates(maker, steps):
    desc = maker()
    s = parity.run_gates(desc, steps=steps)
    assert s["steps"] == steps
    assert s["manifolds"] > 0, "scene produced no contacts: the test checks nothing"
    assert s["worst_manifold"] <= parity.TOL

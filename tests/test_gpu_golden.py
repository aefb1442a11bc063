"""-m gpu: device narrowphase against the committed golden vectors (generated from the reference by
tests/golden/make_golden.py) -- these do not need the oracle at run time."""
import os

import numpy as np
import pytest

from physecs_b200 import scenes as S
from physecs_b200.capi import Context
from tests import parity

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    g = np.load(os.path.join(GOLD, name))
    return {k: g[k] for k in g.files}


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_primitive_manifolds_match_golden():
    d = _golden_module().soup()
    ctx = Context(d)
    ctx.step()
    gm = ctx.manifolds()
    n, worst = parity.compare_manifolds(gm, _gold("narrowphase_prims.npz"), tol=0.0)
    assert n > 100 and worst == 0.0
    ctx.close()


def test_mesh_manifolds_match_golden():
    d = S.terrain(400, cells=24, drop=-0.15)
    ctx = Context(d)
    ctx.step()
    gm = ctx.manifolds()
    # body-body contacts also exist in this scene; the golden holds the body-vs-mesh ones
    sel = gm["keys"][:, 0] == 0
    gm = {k: v[sel] for k, v in gm.items()}
    n, worst = parity.compare_manifolds(gm, _gold("narrowphase_mesh.npz"), tol=0.0)
    assert n > 300 and worst == 0.0
    ctx.close()


def test_convex_manifolds_match_golden():
    """GJK / EPA bin (sphere / capsule / box / convex vs convex) against the reference's physecs::collision."""
    d = _golden_module().convex_soup()
    ctx = Context(d)
    ctx.step()
    gm = ctx.manifolds()
    n, worst = parity.compare_manifolds(gm, _gold("narrowphase_convex.npz"), tol=0.0)
    assert n > 100 and worst == 0.0
    ctx.close()

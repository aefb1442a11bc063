"""-m "not gpu": the oracle (reference compiled by oracle/build_ref.py) against the committed golden vectors, the host-side
setup code against the reference, and the C ABI surface.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from physecs_b200 import scenes as S
from physecs_b200 import capi
from oracle import ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_oracle = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@needs_oracle
def test_oracle_reproduces_golden_narrowphase():
    mg = _golden_module()
    d = mg.soup()
    r = R.RefScene(d, 0, hashfix=True)
    m = r.narrowphase(mg.all_pairs(d.n))
    g = np.load(os.path.join(GOLD, "narrowphase_prims.npz"))
    for k in ("keys", "num_points", "normal", "points"):
        assert np.array_equal(m[k], g[k]), k
    r.close()


@needs_oracle
def test_oracle_reproduces_golden_convex():
    mg = _golden_module()
    d = mg.convex_soup()
    r = R.RefScene(d, 0, hashfix=True)
    m = r.narrowphase(mg.all_pairs(d.n))
    g = np.load(os.path.join(GOLD, "narrowphase_convex.npz"))
    for k in ("keys", "num_points", "normal", "points"):
        assert np.array_equal(m[k], g[k]), k
    r.close()


@needs_oracle
def test_oracle_reproduces_golden_mesh():
    d = S.terrain(400, cells=24, drop=-0.15)
    r = R.RefScene(d, 0, hashfix=True)
    pr = np.stack([np.zeros(400, np.int32), np.zeros(400, np.int32), np.arange(1, 401, dtype=np.int32), np.zeros(400, np.int32)], 1)
    m = r.narrowphase(pr)
    g = np.load(os.path.join(GOLD, "narrowphase_mesh.npz"))
    for k in ("keys", "num_points", "normal", "points"):
        assert np.array_equal(m[k], g[k]), k
    r.close()


@needs_oracle
@pytest.mark.parametrize("hashfix", [False, True])
def test_oracle_reproduces_golden_pyramid(hashfix):
    """As-shipped and hash-fixed builds give bit-identical trajectories (the hash fix only changes speed)."""
    g = np.load(os.path.join(GOLD, "pyramid_steps.npz"))
    r = R.RefScene(S.pyramid(60), 0, hashfix=hashfix)
    for k in range(5):
        r.simulate()
        assert np.array_equal(np.concatenate(r.get_state(), 1), g["states"][k]), f"step {k}"
        assert len(r.pairs()) == g["npairs"][k]
    r.close()


@needs_oracle
def test_oracle_order_hook_is_identity_for_own_order():
    """Feeding the oracle its own manifold order must not change the result (hook plumbing)."""
    d = S.pyramid(40)
    a, b = R.RefScene(d, 0), R.RefScene(d, 0)
    for _ in range(3):
        a.simulate()
        b.simulate()
    a.simulate()
    keys = a.manifold_keys()
    b.set_manifold_order(keys)
    b.simulate()
    assert b.order_stats() == (len(keys), 0, 0)
    for x, y in zip(a.get_state(), b.get_state()):
        assert np.array_equal(x, y)
    # a reversed order changes Gauss-Seidel results: the hook really reorders
    c = R.RefScene(d, 0)
    for _ in range(3):
        c.simulate()
    c.set_manifold_order(keys[::-1].copy())
    c.simulate()
    assert not np.array_equal(c.get_state()[2], a.get_state()[2])
    for s in (a, b, c):
        s.close()


@needs_oracle
@pytest.mark.parametrize("cells", [8, 40])
def test_trimesh_bvh_build_matches_reference(cells):
    """Host-side registration-time BVH build == reference TriangleMesh ctor (triangle order, node bounds, ranges)."""
    m = S.terrain_mesh(cells)
    tri, orig, nb, ci = capi.build_trimesh(m.verts, m.indices)
    d = S.terrain(4, cells=cells)
    r = R.RefScene(d, 0)
    rt, rn, rnb, rci = r.trimesh(0)
    assert np.array_equal(tri, rt) and np.array_equal(nb, rnb) and np.array_equal(ci, rci)
    # tri_orig maps back to the input triangles
    assert np.array_equal(m.indices.reshape(-1, 3)[orig], tri)
    r.close()


def test_trimesh_build_degenerate_inputs():
    # a single triangle and a flat (zero-extent axis) strip build without splitting errors
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1], [1, 0, 1]], np.float32)
    tri, orig, nb, ci = capi.build_trimesh(v, np.array([0, 2, 1], np.uint32))
    assert len(ci) == 1 and ci[0, 0] == 1
    tri, orig, nb, ci = capi.build_trimesh(v, np.array([0, 2, 1, 1, 2, 3], np.uint32))
    assert ci[0, 0] in (0, 2) and sorted(orig.tolist()) == [0, 1]


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "physecs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pb_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = capi.load_library()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in include/physecs_b200.h but not exported: {missing}"
    assert set(capi.EXPORTS) <= declared


def test_no_device_means_failure_not_fallback():
    """Without a CUDA device the context must refuse to exist (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.load_library()
    caps = capi.Caps(16, 16, 64, 64, 4)
    ctx = C.c_void_p()
    assert lib.pb_ctx_create(0, C.byref(caps), C.byref(ctx)) == capi.PB_ECUDA
    with pytest.raises(capi.PbError):
        capi.Context(S.pyramid(4))


def test_product_does_not_import_oracle():
    """The product package never references oracle/ (checker isolation), nor the recording double of the C ABI under tests/."""
    import ast
    pkg = os.path.join(ROOT, "physecs_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            path = os.path.join(dp, fn)
            if fn.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f"{path} imports the oracle"
                    assert not any(n == "tests" or n.startswith("tests.") for n in names), f"{path} imports test infrastructure"
                    if isinstance(node, ast.Constant) and isinstance(node.value, str) and node is not getattr(tree.body[0], "value", None):
                        assert "_ref/" not in node.value and "libphysecs_ref" not in node.value, f"{path} names an oracle artefact"
                        assert "pb_recorder" not in node.value and "libphysecs_b200_scene_recorder" not in node.value, f"{path} names the test double of the C ABI"
            elif fn.endswith((".cu", ".cuh", ".h", ".cpp")):
                for line in open(path, errors="replace"):
                    code = line.split("//")[0]
                    assert not ("#include" in code and "oracle" in code), f"{path} includes oracle code"
                    assert "libphysecs_ref" not in code, f"{path} names an oracle artefact"
                    assert "pb_recorder" not in code and "abi_recorder" not in code, f"{path} names the test double of the C ABI"


def test_scene_generators_are_deterministic():
    a, b = S.terrain(500, cells=32), S.terrain(500, cells=32)
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.col_params, b.col_params) and np.array_equal(a.trimesh[0].verts, b.trimesh[0].verts)
    c = S.mixed_bin(300)
    assert c.n == 305 and c.n_dynamic == 300
    p = S.pyramid(100)
    assert p.n_dynamic == 100 and np.all(p.inv_inertia[1, [0, 4, 8]] > 0)

"""-m "not gpu": the host C++ layer (physecs::Scene mirror, physecs_b200/host/) as far as it runs without a device:
library surface, setup-time helpers against the oracle / the Python restatement, joint colouring, loud failure without a GPU."""
import numpy as np
import pytest

from physecs_b200 import scenes as S
from physecs_b200 import scene_api
from physecs_b200.joint_colors import color_joints
from oracle import ref as R

needs_host = pytest.mark.skipif(not scene_api.available(), reason="host layer not built (needs EnTT / GLM headers)")
needs_oracle = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


@needs_host
def test_scene_library_exports():
    lib = scene_api.load_library()
    missing = [s for s in scene_api.EXPORTS if not hasattr(lib, s)]
    assert not missing, missing


@needs_host
def test_scene_without_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    hs = scene_api.HostScene(S.pyramid(8))
    with pytest.raises(scene_api.SceneError, match="no CPU fallback"):
        hs.simulate()
    hs.close()


@needs_host
@needs_oracle
def test_joint_colouring_matches_reference():
    """Scene::addJoint colours == the reference's JointGraph colours (incl. the overflow index 8, quirk Q21)."""
    for d in (S.joint_star(14), S.ragdolls(3), S.joint_zoo()):
        hs = scene_api.HostScene(d)
        r = R.RefScene(d, 0)
        assert hs.joint_colors == r.joint_colors == color_joints([(j[1], j[4]) for j in d.joints]).tolist()
        hs.close(); r.close()


@needs_host
def test_mass_util_matches_python_restatement():
    """physecs::computeCOMAndInvInertiaTensor (host C++) vs scenes.compute_mass_props (the generator every scene uses):
    compound body with all four solid shape kinds + a trigger collider that must carry no mass."""
    b = S.SceneBuilder("mass")
    meshes = S.convex_templates()
    for m in meshes:
        b.add_convex(m)
    q = S.axis_angle((0.3, 1.0, 0.2), 0.7)
    cols = [dict(type=S.SPHERE, params=(0.3,), lpos=(0.2, 0.0, 0.1)),
            dict(type=S.CAPSULE, params=(0.4, 0.15), lpos=(-0.3, 0.2, 0.0), lquat=tuple(q)),
            dict(type=S.BOX, params=(0.2, 0.3, 0.1), lpos=(0.0, -0.4, 0.2), lquat=tuple(S.axis_angle((1, 0, 0), 0.4))),
            dict(type=S.CONVEX_MESH, params=(0.4, 0.5, 0.3), mesh=3, lpos=(0.1, 0.3, -0.2)),
            dict(type=S.SPHERE, params=(2.0,), flags=S.COL_ENABLE_SIM | S.COL_TRIGGER)]
    e = b.add_body((0, 1, 0), colliders=cols, mass=3.5)
    single = [b.add_body((3 * i, 1, 0), colliders=[c], mass=2.0) for i, c in enumerate(cols[:4], 1)]
    d = b.build()
    hs = scene_api.HostScene(d)
    for ent, mass in [(e, 3.5)] + [(s, 2.0) for s in single]:
        com, inv = hs.mass_props(ent, mass)
        assert np.allclose(com, d.com[ent], rtol=1e-5, atol=1e-6), (ent, com, d.com[ent])
        assert np.allclose(inv, d.inv_inertia[ent], rtol=2e-4, atol=1e-5), (ent, inv, d.inv_inertia[ent])
    hs.close()


@needs_host
def test_trimesh_host_object_matches_device_build():
    """physecs::TriangleMesh built by the host layer has the reference's post-build triangle order (checked against
    pb_build_trimesh, which test_cpu_oracle pins to the reference's TriangleMesh ctor)."""
    from physecs_b200 import capi
    m = S.terrain_mesh(12)
    tri, orig, nb, ci = capi.build_trimesh(m.verts, m.indices)
    assert len(ci) >= 1 and tri.shape == (len(m.indices) // 3, 3)
    d = S.terrain(4, cells=12)
    hs = scene_api.HostScene(d)     # constructs a physecs::TriangleMesh from the same vertices / indices
    hs.close()

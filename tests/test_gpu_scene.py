"""-m gpu: the host C++ layer (physecs::Scene over an entt::registry, include/Physecs/Physecs.h of this repo) against the
oracle (the reference's physecs::Scene compiled by oracle/build_ref.py), both driven through the same public API:
registry components in, Scene::simulate, registry components out.  The oracle is fed the device's colour-batched contact
order each step (north_star gate 3); the two registries must then stay IDENTICAL while both free-run."""
import numpy as np
import pytest

from physecs_b200 import scenes as S
from physecs_b200 import scene_api
from tests import parity

pytestmark = pytest.mark.gpu


def _events(ev):
    return sorted(map(tuple, np.asarray(ev).tolist()))


def run_scene(desc, steps, ops=None, threads=2, contact_filter=0, device_authoritative=False, teacher=False, tol=0.0, arena=None):
    from oracle.ref import RefScene
    ref = RefScene(desc, 0, hashfix=True)
    hs = scene_api.HostScene(desc, num_threads=threads)
    if arena:
        hs.set_arena_capacity(*arena)
    if contact_filter:
        ref.set_contact_filter(contact_filter); hs.set_contact_filter(contact_filter)
    if device_authoritative:
        hs.set_sync_mode(True)
    ref.record_trigger_events(); hs.record_trigger_events()
    out = dict(manifolds=0, pairs=0, triggers=0, events=0, worst=0.0)
    try:
        for k in range(steps):
            if ops and k in ops:
                ops[k](ref); ops[k](hs)
            if teacher:
                p, q, v, w = ref.get_state()
                hs.set_state(np.arange(len(p)), p, q, v, w)
            hs.simulate()
            taps = hs.taps()
            gm = taps.manifolds()
            ref.set_manifold_order(gm["keys"])
            ref.simulate()
            matched, missing, extra = ref.order_stats()
            assert missing == 0 and extra == 0, f"step {k}: manifold sets differ: matched={matched} missing={missing} extra={extra}"
            out["pairs"] = max(out["pairs"], parity.compare_pairs(taps.pairs(), ref.pairs()))
            out["triggers"] = max(out["triggers"], parity.compare_triggers(taps.triggers(), ref.triggers()))
            ge, re_ = _events(hs.take_trigger_events()), _events(ref.take_trigger_events())
            assert ge == re_, f"step {k}: trigger listener calls differ: {ge[:4]} vs {re_[:4]}"
            out["events"] += len(ge)
            P, Q, V, W = hs.get_state()
            p, q, v, w = ref.get_state()
            sgn = np.sign(np.sum(Q * q, axis=1, keepdims=True)); sgn[sgn == 0] = 1
            errs = dict(pos=parity.max_err(P, p), quat=parity.max_err(Q * sgn, q), vel=parity.max_err(V, v), angvel=parity.max_err(W, w))
            for name, e in errs.items():
                assert e <= tol, f"step {k}: registry state differs from the oracle: {errs} ({matched} manifolds)"
                out["worst"] = max(out["worst"], e)
            out["manifolds"] = max(out["manifolds"], matched)
        out["stats"] = hs.stats()
    finally:
        hs.close(); ref.close()
    return out


@pytest.mark.parametrize("maker,steps", [
    (lambda: S.pyramid(120), 40),
    (lambda: S.mixed_bin(800, spacing=0.8), 60),
    (lambda: S.terrain_mixed(500, cells=32), 70),
    (lambda: S.ragdolls(6), 80),
    (lambda: S.convex_pile(200, mix_prims=True), 60),
])
def test_scene_free_run_identical_to_oracle(maker, steps):
    r = run_scene(maker(), steps)
    assert r["worst"] == 0.0 and r["manifolds"] > 0


def test_scene_joint_zoo_and_gears_teacher_forced():
    """Servo (acosf) and gear (atan2f) rows are at the 1e-4 gate, so these scenes are compared step by step from the oracle's state."""
    for d in (S.joint_zoo(), S.gear_train(3), S.joint_star(12)):
        r = run_scene(d, 50, teacher=True, tol=parity.TOL)
        assert r["worst"] <= parity.TOL


@pytest.mark.parametrize("contact_filter", [0, 2])
def test_scene_trigger_listeners(contact_filter):
    r = run_scene(S.trigger_zoo(120), 70, contact_filter=contact_filter)
    assert r["worst"] == 0.0 and r["triggers"] > 10 and r["events"] > 40


def test_scene_structural_edits():
    """Bodies spawned and destroyed mid-run, kinematic toggles, a patched static wall, setCanCollide: the device scene is
    re-uploaded behind the API and must stay on the oracle's trajectory (bounds history + contact cache carried over)."""
    d = S.mixed_bin(500, spacing=0.8)
    extra = S.mixed_bin(60, spacing=0.8, seed=0x99)
    sel = extra.dynamic_entities()
    import dataclasses
    cut = lambda a: np.ascontiguousarray(a[sel])
    spawn = dataclasses.replace(extra, pos=cut(extra.pos) + np.array([0, 6.0, 0], np.float32), quat=cut(extra.quat), flags=cut(extra.flags), vel=cut(extra.vel),
                                angvel=cut(extra.angvel), inv_mass=cut(extra.inv_mass), com=cut(extra.com), inv_inertia=cut(extra.inv_inertia),
                                col_offsets=np.arange(len(sel) + 1, dtype=np.int32), col_lpos=cut(extra.col_lpos), col_lquat=cut(extra.col_lquat),
                                col_type=cut(extra.col_type), col_params=cut(extra.col_params), col_mesh=cut(extra.col_mesh),
                                col_material=cut(extra.col_material), col_flags=cut(extra.col_flags), col_data=cut(extra.col_data))
    assert np.array_equal(extra.col_offsets, np.arange(extra.n + 1))
    dyn = d.dynamic_entities()
    wall = int(d.static_entities()[1])
    ops = {
        12: lambda s: s.add_entities(spawn),
        20: lambda s: [s.destroy_entity(int(e)) for e in dyn[[3, 40, 77, 120, 300]]],
        24: lambda s: s.set_kinematic(int(dyn[10]), True),
        28: lambda s: s.set_state([wall], d.pos[[wall]] + np.array([[0.05, 0, 0]], np.float32), d.quat[[wall]], patch=True),
        32: lambda s: [s.set_can_collide(int(dyn[5]), int(dyn[6]), False), s.set_can_collide(int(dyn[50]), int(d.static_entities()[0]), False)],
        36: lambda s: s.set_kinematic(int(dyn[10]), False),
        40: lambda s: s.add_entities(spawn),
    }
    r = run_scene(d, 60, ops=ops)
    assert r["worst"] == 0.0 and r["manifolds"] > 100


def test_scene_collider_edits():
    """Scene::addCollider / clearColliders on live entities (Physecs.cpp:725-751): a second collider on dynamic bodies and on a
    static wall, bodies left without colliders -- the device collider table is re-uploaded, everything that persists keeps its
    bounds history and contact cache, the new colliders start from creation bounds."""
    d = S.mixed_bin(400, spacing=0.8)
    dyn = d.dynamic_entities()
    floor = int(d.static_entities()[0])
    idq = [0, 0, 0, 1]
    bouncy = (0.4, 0.3, 0.0)
    ops = {
        8: lambda s: [s.add_collider(int(e), [0.45, 0.0, 0.0], idq, S.SPHERE, [0.25], material=bouncy) for e in dyn[[7, 8, 150]]],
        14: lambda s: s.add_collider(floor, [0.0, 1.3, 0.0], idq, S.BOX, [1.5, 0.3, 1.5], material=bouncy),
        20: lambda s: [s.clear_colliders(int(e)) for e in dyn[[30, 31]]],
        26: lambda s: s.add_collider(int(dyn[30]), [0.0, 0.0, 0.0], idq, S.CAPSULE, [0.3, 0.2]),
        34: lambda s: s.add_collider(int(dyn[9]), [0.0, 0.4, 0.0], [0.0, 0.0, 0.38268343, 0.92387953], S.BOX, [0.3, 0.1, 0.2], material=bouncy),
    }
    r = run_scene(d, 50, ops=ops)
    assert r["worst"] == 0.0 and r["manifolds"] > 100


def test_scene_collider_cleared_and_added_again():
    """A collider cleared and added again under the same (entity, index) between two steps: new broadphase entry from creation
    bounds, but the reference's contact cache goes by name and still finds the pair's entries (Physecs.cpp:237)."""
    d = S.mixed_bin(300, spacing=0.8)
    d.col_material[:, 1] = 0.3            # restitution: the cache's target velocities matter
    dyn = d.dynamic_entities()
    idq = [0, 0, 0, 1]

    def again(s, e):
        s.clear_colliders(e)
        s.add_collider(e, d.col_lpos[e], d.col_lquat[e], int(d.col_type[e]), list(d.col_params[e]), material=tuple(d.col_material[e]))
    ops = {k: (lambda s, k=k: [again(s, int(e)) for e in dyn[[k, k + 50, k + 100]]]) for k in (22, 23, 24, 25, 26, 30, 34)}
    r = run_scene(d, 45, ops=ops)
    assert r["worst"] == 0.0 and r["manifolds"] > 100


def test_scene_dynamic_pool_reordered():
    """registry.sort on the RigidBodyDynamicComponent pool between two steps: no signal fires, the reference simply indexes bodies by
    their new position (Physecs.cpp:116-117) -- the Scene has to notice that its rows no longer follow the pool and re-upload (joints
    and contacts included: ragdolls on the ground).  EnTT iterates a pool back to front, so the comparator `a > b` (True) leaves a pool
    filled in creation order as it is and `a < b` (False) reverses its packed order: step 18 is the no-op, step 30 the reversal."""
    d = S.ragdolls(5)
    ops = {18: lambda s: s.sort_dynamic(True), 30: lambda s: s.sort_dynamic(False)}
    r = run_scene(d, 44, ops=ops)
    assert r["worst"] == 0.0 and r["manifolds"] > 0


def test_scene_joint_edits():
    d = S.ragdolls(4)
    j0 = d.joints[3]
    ops = {
        15: lambda s: s.destroy_joint(3),
        25: lambda s: s.add_joint(*j0),
        30: lambda s: s.set_revolute_drive(3 + 10, True, 2.0, 3.0),     # joint 13 = the same elbow of ragdoll 1 (revolute)
        45: lambda s: s.set_revolute_drive(3 + 10, False, 0.0, 0.0),
    }
    assert d.joints[13][0] == S.J_REVOLUTE
    r = run_scene(d, 60, ops=ops)
    assert r["worst"] == 0.0


def test_scene_device_authoritative_mode():
    """SYNC_DEVICE_AUTHORITATIVE skips the per-step registry read; bodies the application writes are announced with patch."""
    d = S.mixed_bin(400, spacing=0.8)
    dyn = d.dynamic_entities()
    tele = lambda s: s.set_state([int(dyn[7])], np.array([[0.0, 9.0, 0.0]], np.float32), np.array([[0, 0, 0, 1]], np.float32),
                                 np.array([[0, -1.0, 0]], np.float32), np.zeros((1, 3), np.float32), patch=True)
    r = run_scene(d, 50, ops={20: tele}, device_authoritative=True)
    assert r["worst"] == 0.0


def test_scene_arena_growth():
    """Arenas that are far too small: the overflowing step is re-run with larger arenas and nothing diverges."""
    r = run_scene(S.mixed_bin(600, spacing=0.8), 40, arena=(64, 64))
    assert r["worst"] == 0.0 and r["pairs"] > 64

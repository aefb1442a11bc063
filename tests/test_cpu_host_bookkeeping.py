"""-m "not gpu": the bookkeeping of the host layer (physecs_b200/host/Scene.cpp) -- row <-> entity maps, the collider table, what a
structural edit carries over, where the read-back lands -- checked without a device.  The host layer is built UNCHANGED against
tests/abi_recorder/pb_recorder.cpp, a recording double of the C ABI that computes nothing (its "step" adds 1 to every non-kinematic
dynamic row's pos.x; collider bounds are opaque (upload number, collider index) tags).  The arithmetic of a step is the GPU tests'
business (tests/test_gpu_scene.py runs the same edits against the reference on a B200); what is checked here is that every row and
collider the device is told about stands for the right entity, before and after edits (reference semantics: Physecs.cpp:20-77,
:116-117, :725-751)."""
import ctypes as C

import numpy as np
import pytest

from physecs_b200 import scenes as S
from physecs_b200 import scene_api
from tests.abi_recorder import build as recorder_build

_built = recorder_build.build()
pytestmark = pytest.mark.skipif(_built is None, reason="EnTT / GLM headers not available: the host layer cannot be built here")


class Recorded:
    """HostScene over the recording double + readers of what the host layer told the 'device'."""

    def __init__(self, desc, threads=2):
        rec_path, scene_path = _built
        self.rec = C.CDLL(rec_path)
        self.hs = scene_api.HostScene(desc, num_threads=threads, lib=scene_api.load_other_build(scene_path))
        self.hs.simulate()

    @property
    def ctx(self):
        return C.c_void_p(self.hs.lib.psh_native_context(self.hs.h))

    def counts(self):
        out = (C.c_int * 8)()
        self.rec.pbr_counts(self.ctx, out)
        return dict(zip(("n_dyn", "n_static", "n_col", "uploads", "steps", "joints", "no_collide", "filter_classes"), list(out)))

    def rows(self):
        c = self.counts()
        e = np.zeros(c["n_dyn"] + c["n_static"], np.int32)
        self.rec.pbr_rows(self.ctx, e.ctypes.data_as(C.POINTER(C.c_int)))
        return e

    def colliders(self):
        n = self.counts()["n_col"]
        row, idx, typ = (np.zeros(n, np.int32) for _ in range(3))
        tag = np.zeros((n, 2), np.float32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self.rec.pbr_colliders(self.ctx, ip(row), ip(idx), ip(typ), tag.ctypes.data_as(C.POINTER(C.c_float)))
        return row, idx, typ, tag

    def named_tags(self):
        """{(entity, collider index): (upload number, collider position at that upload)}"""
        ent = self.rows()
        row, idx, _, tag = self.colliders()
        return {(int(ent[r]), int(i)): (int(t[0]), int(t[1])) for r, i, t in zip(row, idx, tag)}

    def carry_map(self, which):
        out = (C.c_int * (1 << 16))()
        n = self.rec.pbr_map(self.ctx, which, 1 << 16, out)
        return np.array(out[:n], np.int32)

    def joints(self):
        n = self.counts()["joints"]
        r0, r1, col = (np.zeros(n, np.int32) for _ in range(3))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self.rec.pbr_joints(self.ctx, ip(r0), ip(r1), ip(col))
        return r0, r1, col

    def close(self):
        self.hs.close()


def _moved(now, start, k):
    """x advanced by k unit "steps" (added one at a time in fp32, so compare with a tolerance), y / z untouched"""
    return bool(np.all(np.abs(now[:, 0] - (start[:, 0] + np.float32(k))) < 1e-4)) and np.array_equal(now[:, 1:], start[:, 1:])


def _check_tables(r, alive_entities, desc_of):
    """Rows: dynamic entities first, each alive rigid body exactly once; colliders row-major, indices 0..k-1 per body, types as described."""
    c = r.counts()
    ent = r.rows()
    assert sorted(ent.tolist()) == sorted(alive_entities)
    row, idx, typ, _ = r.colliders()
    assert np.all(np.diff(row) >= 0), "colliders are not row-major"
    for rw in np.unique(row):
        k = idx[row == rw]
        assert k.tolist() == list(range(len(k)))
    for rw, i, t in zip(row, idx, typ):
        assert desc_of(int(ent[rw]))[int(i)] == int(t)
    return c


def test_rows_and_colliders_follow_the_registry():
    d = S.mixed_bin(300, spacing=0.8)
    r = Recorded(d)
    types = {e: [int(t) for t in d.col_type[d.col_offsets[e]:d.col_offsets[e + 1]]] for e in range(d.n)}
    c = _check_tables(r, list(range(d.n)), lambda e: types[e])
    assert (c["n_dyn"], c["n_static"], c["n_col"], c["uploads"], c["steps"]) == (d.n_dynamic, d.n - d.n_dynamic, len(d.col_type), 1, 1)
    ent = r.rows()
    assert set(ent[:c["n_dyn"]].tolist()) == set(d.dynamic_entities().tolist())
    # the read-back lands on the entities the rows stand for: every dynamic body moved by exactly one "step"
    for k in range(2, 5):
        r.hs.simulate()
        p = r.hs.get_state()[0]
        dyn = d.dynamic_entities()
        assert _moved(p[dyn], d.pos[dyn], k)
        st = d.static_entities()
        assert np.array_equal(p[st], d.pos[st])
    r.close()


def test_spawn_and_destroy_carry_bounds_and_cache_by_name():
    d = S.mixed_bin(200, spacing=0.8)
    r = Recorded(d)
    before = r.named_tags()
    assert all(t[0] == 1 for t in before.values())
    dyn = d.dynamic_entities()
    gone = [int(e) for e in dyn[[3, 40, 77]]]
    for e in gone:
        r.hs.destroy_entity(e)
    r.hs.simulate()
    after = r.named_tags()
    assert set(after) == {k for k in before if k[0] not in gone}
    assert all(after[k] == before[k] for k in after), "a surviving collider lost its bounds history"
    c = r.counts()
    assert (c["n_dyn"], c["uploads"]) == (d.n_dynamic - 3, 2)
    # the maps the device got: old collider -> new collider, -1 for the destroyed ones, the same for bounds and contact cache
    bm, cm = r.carry_map(0), r.carry_map(1)
    assert len(bm) == len(before) and np.array_equal(bm, cm) and int(np.count_nonzero(bm < 0)) == 3
    assert sorted(bm[bm >= 0].tolist()) == list(range(c["n_col"]))
    # spawn: EnTT recycles the destroyed entities' indices under a new version -- new bodies must not inherit anything
    extra = S.dynamic_only(S.mixed_bin(5, spacing=0.8, seed=0x99), lift=(0.0, 6.0, 0.0))
    first = r.hs.add_entities(extra)
    r.hs.simulate()
    now = r.named_tags()
    fresh = {k: v for k, v in now.items() if k not in after}
    assert len(fresh) == 5 and all(v[0] == 3 for v in fresh.values())
    assert all(now[k] == after[k] for k in after)
    bm, cm = r.carry_map(0), r.carry_map(1)
    assert np.array_equal(bm, cm) and int(np.count_nonzero(bm < 0)) == 0 and len(bm) == len(after)
    # positions: the survivors have taken 3 steps, the newcomers 1
    p = r.hs.get_state()[0]
    keep = np.array([e for e in dyn if int(e) not in gone])
    assert _moved(p[keep], d.pos[keep], 3)
    assert _moved(p[first:first + 5], extra.pos, 1)
    r.close()


def test_collider_edits():
    d = S.mixed_bin(60, spacing=0.8)
    r = Recorded(d)
    dyn = d.dynamic_entities()
    a, b = int(dyn[7]), int(dyn[9])
    base = r.named_tags()
    idq = [0, 0, 0, 1]
    r.hs.add_collider(a, [0.4, 0, 0], idq, S.SPHERE, [0.2])               # a second collider: new name, creation bounds
    r.hs.clear_colliders(b)                                               # cleared and added again: old name ...
    r.hs.add_collider(b, d.col_lpos[b], d.col_lquat[b], int(d.col_type[b]), list(d.col_params[b]))
    r.hs.simulate()
    now = r.named_tags()
    assert set(now) == set(base) | {(a, 1)}
    assert now[(a, 1)][0] == 2 and now[(b, 0)][0] == 2, "a collider created since the last upload must start from creation bounds"
    assert all(now[k] == base[k] for k in base if k != (b, 0))
    bm, cm = r.carry_map(0), r.carry_map(1)
    old_b = base[(b, 0)][1]
    assert bm[old_b] == -1 and cm[old_b] >= 0, "the contact cache goes by name (Physecs.cpp:237), the broadphase entry is new (Physecs.cpp:739-748)"
    assert int(np.count_nonzero(bm < 0)) == 1 and int(np.count_nonzero(cm < 0)) == 0
    _, _, typ, _ = r.colliders()
    assert r.counts()["n_col"] == len(base) + 1 and int(np.count_nonzero(typ == S.SPHERE)) == int(np.count_nonzero(d.col_type == S.SPHERE)) + 1
    r.close()


def test_reordered_dynamic_pool_is_noticed():
    d = S.ragdolls(3)
    r = Recorded(d)
    rows0 = r.rows().copy()
    j0 = r.joints()
    n_dyn = r.counts()["n_dyn"]
    r.hs.simulate()
    assert r.counts()["uploads"] == 1
    r.hs.sort_dynamic(True)                  # comparator a > b: EnTT iterates back to front, the packed order stays as it is
    r.hs.simulate()
    assert r.counts()["uploads"] == 1
    r.hs.sort_dynamic(False)                 # the packed order is reversed; no signal fires
    r.hs.simulate()
    c = r.counts()
    rows1 = r.rows()
    assert c["uploads"] == 2 and rows1[:n_dyn].tolist() == rows0[:n_dyn].tolist()[::-1]
    # joints name the same ENTITIES through the new rows, colours unchanged
    j1 = r.joints()
    assert np.array_equal(rows0[j0[0]], rows1[j1[0]]) and np.array_equal(rows0[j0[1]], rows1[j1[1]]) and np.array_equal(j0[2], j1[2])
    assert np.array_equal(r.carry_map(2), np.arange(c["joints"])), "joints already on the device keep their state"
    # every collider kept its history, and every body has taken exactly 4 steps
    assert all(t[0] == 1 for t in r.named_tags().values())
    p = r.hs.get_state()[0]
    dyn = d.dynamic_entities()
    assert _moved(p[dyn], d.pos[dyn], 4)
    r.hs.simulate()
    assert r.counts()["uploads"] == 2        # ... and nothing is re-uploaded while the pool stays as it is
    r.close()


def test_patched_transform_moves_the_row_and_marks_its_bounds():
    d = S.mixed_bin(40, spacing=0.8)
    r = Recorded(d)
    wall = int(d.static_entities()[1])
    newp = d.pos[[wall]] + np.array([[0.25, 0, 0]], np.float32)
    r.hs.set_state([wall], newp, d.quat[[wall]], patch=True)
    r.hs.set_state([wall], newp + np.float32(0.5), d.quat[[wall]], patch=True)          # patched twice: the last one wins, one row goes out
    r.hs.simulate()
    tags = r.named_tags()
    assert tags[(wall, 0)][0] == -1 and sum(1 for t in tags.values() if t[0] == -1) == 1
    assert r.counts()["uploads"] == 1
    r.close()


def test_kinematic_body_is_not_written_back():
    d = S.mixed_bin(30, spacing=0.8)
    r = Recorded(d)
    e = int(d.dynamic_entities()[4])
    r.hs.set_kinematic(e, True)
    r.hs.simulate()
    p = r.hs.get_state()[0]
    assert _moved(p[[e]], d.pos[[e]], 1)       # one step before the flag, none after (Physecs.cpp:446, :497)
    others = np.array([x for x in d.dynamic_entities() if x != e])
    assert _moved(p[others], d.pos[others], 2)
    r.close()


def test_arena_overflow_grows_the_arenas_and_runs_the_step_again():
    """pb_step only enqueues; an overflowing arena is reported by the read-back (PB_ECAPACITY) with the counters saying what the step needs.
    Scene::simulate enlarges the arenas in place and runs the step again -- once, not once per missing entry."""
    d = S.mixed_bin(50, spacing=0.8)
    r = Recorded(d)
    caps = (C.c_int * 5)()
    r.rec.pbr_caps(r.ctx, caps)
    pairs0, man0 = caps[2], caps[3]
    r.rec.pbr_fail_next_steps(r.ctx, 5, pairs0 * 3, man0 + 10)
    r.hs.simulate()
    r.rec.pbr_caps(r.ctx, caps)
    assert caps[2] >= pairs0 * 3 and caps[3] >= man0 + 10
    c = r.counts()
    assert c["steps"] == 2 and c["uploads"] == 1            # the first step and the repeated one; the scene was not re-uploaded
    p = r.hs.get_state()[0]
    dyn = d.dynamic_entities()
    assert _moved(p[dyn], d.pos[dyn], 2)                    # the failed attempt moved nothing
    r.close()


def test_a_failure_no_arena_explains_is_reported_not_retried():
    d = S.mixed_bin(20, spacing=0.8)
    r = Recorded(d)
    r.rec.pbr_fail_next_steps(r.ctx, 100, -1, -1)           # PB_ECAPACITY with the counters inside the arenas: nothing to grow
    with pytest.raises(scene_api.SceneError, match="recorded"):
        r.hs.simulate()
    assert r.counts()["steps"] == 1                         # one attempt, no blind retries
    r.rec.pbr_fail_next_steps(r.ctx, 0, 0, 0)
    r.hs.simulate()                                         # ... and the Scene is usable afterwards
    assert r.counts()["steps"] == 2
    r.close()


def test_contact_filter_table_and_noncolliding_pairs():
    d = S.trigger_zoo(60)
    r = Recorded(d)
    assert r.counts()["filter_classes"] == 0                # defaultContactFilter: no table
    r.hs.set_contact_filter(1)
    r.hs.simulate()
    classes = {(int(f) & 1, int(x)) for f, x in zip(d.col_flags, d.col_data)}
    assert r.counts()["filter_classes"] == len(classes)     # one class per (isTrigger, data) present in the scene
    n0 = r.counts()["no_collide"]
    dyn = d.dynamic_entities()
    r.hs.set_can_collide(int(dyn[0]), int(dyn[1]), False)
    r.hs.set_can_collide(int(dyn[1]), int(dyn[0]), False)   # the same pair the other way round
    r.hs.simulate()
    assert r.counts()["no_collide"] == n0 + 1
    r.hs.set_can_collide(int(dyn[0]), int(dyn[1]), True)
    # a spawn re-uploads the colliders: the per-collider class table has to follow
    r.hs.add_entities(S.dynamic_only(S.mixed_bin(3, spacing=0.8, seed=0x99), lift=(0.0, 6.0, 0.0)))
    r.hs.simulate()
    c = r.counts()
    assert c["no_collide"] == n0 and c["filter_classes"] >= len(classes) and c["uploads"] == 2
    r.close()


def test_outgrown_context_is_replaced_and_the_bounds_history_travels_through_the_host():
    """A scene that outgrows max_bodies / max_colliders gets a new device context (capacities carry 50 % headroom): poses, velocities and
    the bounds history are carried over (pb_get_bounds before, pb_set_bounds after), the contact cache starts empty (documented)."""
    d = S.mixed_bin(300, spacing=0.8)
    r = Recorded(d)
    old_ctx = r.ctx.value
    before = r.named_tags()
    row0, idx0, _, _ = r.colliders()
    dyn = d.dynamic_entities()
    gone = [int(e) for e in dyn[[0, 1, 2]]]
    for e in gone:
        r.hs.destroy_entity(e)
    extra = S.dynamic_only(S.mixed_bin(1500, spacing=0.8, seed=0x99), lift=(0.0, 30.0, 0.0))
    first = r.hs.add_entities(extra)
    r.hs.sort_dynamic(False)                                 # ... all of them: the pool is reversed as well
    r.hs.simulate()
    assert r.ctx.value != old_ctx or r.counts()["uploads"] == 1, "the context was expected to be replaced"
    c = r.counts()
    assert c["uploads"] == 1 and c["n_dyn"] == d.n_dynamic - 3 + 1500
    now = r.named_tags()
    survivors = {k for k in before if k[0] not in gone}
    assert survivors <= set(now)
    assert all(now[k] == before[k] for k in survivors), "a surviving collider lost its bounds history across the context change"
    ent = r.rows()
    row, idx, _, tag = r.colliders()
    moved = [i for i in range(len(row)) if (int(ent[row[i]]), int(idx[i])) in survivors and int(tag[i][1]) != i]
    assert len(moved) > 100, "the check above only means something where collider positions changed"
    fresh = [i for i in range(len(row)) if (int(ent[row[i]]), int(idx[i])) not in survivors]
    assert len(fresh) == 1500 and all(int(tag[i][0]) == 1 and int(tag[i][1]) == i for i in fresh)
    assert len(r.carry_map(1)) == 0                          # no contact cache to re-key on a new context
    p = r.hs.get_state()[0]
    keep = np.array([e for e in dyn if int(e) not in gone])
    assert _moved(p[keep], d.pos[keep], 2) and _moved(p[first:first + 1500], extra.pos, 1)
    r.close()

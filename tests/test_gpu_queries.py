"""-m gpu: scene queries of the host C++ layer (Scene::raycastClosest, Scene::overlap, Scene::overlapWithMinTranslationalDistance -- SURVEY.md §8f-3) against the reference's
own Scene on identical registries: before the first step (creation-time bounds) and after stepping (refreshed bounds, moved bodies)."""
import numpy as np
import pytest

from physecs_b200 import scenes as S
from physecs_b200 import scene_api

pytestmark = pytest.mark.gpu


def _step_both(hs, ref, steps):
    for _ in range(steps):
        hs.simulate()
        ref.set_manifold_order(hs.taps().manifolds()["keys"])
        ref.simulate()
        assert ref.order_stats()[1:] == (0, 0)
    for a, b in zip(hs.get_state(), ref.get_state()):
        assert np.array_equal(a, b), "registries diverged: the query comparison would be meaningless"


def _rays(n, seed, centre, extent):
    rng = np.random.default_rng(seed)
    o = centre + (rng.random((n, 3)) - 0.5) * extent
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:, 1] = -np.abs(d[:, 1])            # mostly downwards so many rays hit something
    return o.astype(np.float32), d.astype(np.float32)


def _mtd_rows(res):
    ids, val = res
    rows = [tuple(i) + tuple(v) for i, v in zip(ids.tolist(), val.view(np.int32).tolist())]    # float bits: the comparison is exact
    return sorted(rows)


def _compare_queries(hs, ref, d, seed, min_found=30, centre_y=2.0):
    centre = np.array([0.0, centre_y, 0.0]); extent = np.array([16.0, 4.0, 16.0])
    o, dr = _rays(400, seed, centre, extent)
    hits = ties = 0
    for i in range(len(o)):
        for mod, skip in ((0, 0), (3, 1)):
            e0, p0 = hs.raycast(o[i], dr[i], 30.0, mod, skip)
            e1, p1 = ref.raycast(o[i], dr[i], 30.0, mod, skip)
            assert (e0 >= 0) == (e1 >= 0), f"ray {i} filter {(mod, skip)}: hit {e0} vs reference {e1}"
            if e0 >= 0:
                hits += 1
                assert np.allclose(p0, p1, rtol=0, atol=1e-5), (i, e0, e1, p0, p1)
                # equal distance (a ray starting inside several overlapping shapes, t = 0): which one the reference names
                # depends on the shape of its incremental BVH; the distance is what is defined
                ties += e0 != e1
    assert hits > 100, "too few ray hits: the test checks little"
    assert ties <= 0.1 * hits, f"{ties} of {hits} hits name a different entity at equal distance"
    rng = np.random.default_rng(seed + 1)
    meshes = len(d.convex)
    found = found_mtd = 0
    for k in range(120):
        pos = (centre + (rng.random(3) - 0.5) * extent * np.array([1, 0.6, 1])).astype(np.float32)
        q = rng.normal(size=4); q = (q / np.linalg.norm(q)).astype(np.float32)
        t = k % 4
        prm = [(0.5 + rng.random(),), (0.4 + rng.random() * 0.5, 0.3), (0.4 + rng.random(), 0.3, 0.6), (0.8, 0.7, 0.9)][t]
        mesh = int(rng.integers(meshes)) if (t == S.CONVEX_MESH and meshes) else -1
        if t == S.CONVEX_MESH and not meshes:
            continue
        for flt in (0, 1):
            a = sorted(map(tuple, hs.overlap(pos, q, t, prm, mesh, flt).tolist()))
            b = sorted(map(tuple, ref.overlap(pos, q, t, prm, mesh, flt).tolist()))
            assert a == b, f"overlap query {k} (type {t}, filter {flt}): {a[:4]} vs reference {b[:4]}"
            found += len(a)
        # overlapWithMinTranslationalDistance: one row per manifold of collision(collider, query shape), triangle meshes included
        a = _mtd_rows(hs.overlap_mtd(pos, q, t, prm, mesh))
        b = _mtd_rows(ref.overlap_mtd(pos, q, t, prm, mesh))
        assert a == b, f"overlapWithMinTranslationalDistance query {k} (type {t}): {len(a)} rows {a[:2]} vs reference {len(b)} rows {b[:2]}"
        found_mtd += len(a)
    assert found > min_found, "overlap queries found almost nothing"
    assert found_mtd > min_found or min_found == 0, "overlapWithMinTranslationalDistance queries found almost nothing"


# (scene, minimum number of overlap results the random queries must find: the 4-ragdoll scene is 48 small bodies, so few are met: any is enough)
@pytest.mark.parametrize("maker,min_found,centre_y", [(lambda: S.trigger_zoo(140), 30, 2.0), (lambda: S.convex_pile(200, mix_prims=True), 30, 2.0),
                                                      (lambda: S.ragdolls(4), 0, 2.0), (lambda: S.terrain_mixed(600, cells=40), 30, 0.8)])
def test_raycast_and_overlap_match_reference(maker, min_found, centre_y):
    from oracle.ref import RefScene
    d = maker()
    ref = RefScene(d, 0, hashfix=True)
    hs = scene_api.HostScene(d, num_threads=2)
    try:
        _compare_queries(hs, ref, d, 11, min_found, centre_y)          # before the first simulate: creation-time bounds, no margin
        _step_both(hs, ref, 25)
        _compare_queries(hs, ref, d, 23, min_found, centre_y)          # after stepping: refreshed bounds, moved bodies
        # a static body announced through registry.patch is seen by the next query without a simulate in between
        st = int(d.static_entities()[0])
        for s in (hs, ref):
            s.set_state([st], d.pos[[st]] + np.array([[0.0, 0.4, 0.0]], np.float32), d.quat[[st]], patch=True)
        _compare_queries(hs, ref, d, 37, min_found, centre_y)
    finally:
        hs.close(); ref.close()


def test_bvh_debug_snapshot():
    """Scene::getBVH / getBHVRootId (debug getters, Physecs.h:227-228): a well-formed tree (checked in the harness: parent links,
    parent boxes hold their children, every node reachable from the root) whose leaves are exactly the reference Scene's
    broadphase entries with bit-identical bounds."""
    from oracle.ref import RefScene
    d = S.mixed_bin(400, spacing=0.8)
    ref = RefScene(d, 0, hashfix=True)
    hs = scene_api.HostScene(d, num_threads=0)
    try:
        for phase in range(2):
            ids, b = hs.bvh_leaves()
            rid, rb = ref.bounds()
            assert len(ids) == len(rid) == len(d.col_type)
            o, ro = np.lexsort(ids.T[::-1]), np.lexsort(rid.T[::-1])
            assert np.array_equal(ids[o], rid[ro])
            assert np.array_equal(b[o], rb[ro]), "leaf bounds differ from BroadPhaseEntry::bounds"
            _step_both(hs, ref, 10)
    finally:
        hs.close(); ref.close()

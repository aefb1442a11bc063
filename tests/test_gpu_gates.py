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


def test_joint_overflow_bucket():
    """More than 8 joints on one body: the 9th.. land in the reference's sequential overflow bucket (scalar Constraint1D
    semantics, quirk Q9); the device solves that bucket sequentially with the same arithmetic."""
    from physecs_b200.joint_colors import color_joints
    d = S.joint_star(12)
    colors = color_joints([(j[1], j[4]) for j in d.joints])
    assert (colors == 8).sum() >= 4
    s = parity.run_gates(d, steps=60)
    assert s["steps"] == 60


def test_gear_joint():
    """GearJoint (persistent angle state across steps; atan2f / fmodf -> 1e-4 gate)."""
    s = parity.run_gates(S.gear_train(3), steps=60)
    assert s["steps"] == 60


@pytest.mark.parametrize("contact_filter", [0, 1, 2])
def test_trigger_pairs(contact_filter):
    """physecs::overlap on TRIGGER pairs (every shape combination) + the three gates on the colliding rest; custom contact
    filters (Scene::setContactFilter) go through the tabulated (isTrigger, data) class table."""
    s = parity.run_gates(S.trigger_zoo(160), steps=70, contact_filter=contact_filter)
    assert s["steps"] == 70
    assert s["triggers"] > 20, "no overlapping trigger pairs: the test checks nothing"
    assert s["trigger_changes"] > 40, "trigger sets never changed: enter / exit not exercised"


@pytest.mark.parametrize("local_max", [4, 1024])
@pytest.mark.parametrize("maker,steps", [
    (lambda: S.ragdolls(8), 50), (lambda: S.terrain_mixed(500, cells=32), 60), (lambda: S.mixed_bin(700, spacing=0.8), 50), (lambda: S.joint_star(12), 40),
])
def test_three_gates_with_island_sweeps(maker, steps, local_max, monkeypatch):
    """Simulation islands forced on (default: auto): small islands are swept per CTA, the rest device-wide.  local_max = 4 pushes
    every island with more than four constraints into the device-wide sweep, so both kinds run side by side in one step; the solve
    order handed to the oracle is then (group, colour, slot).  Results must stay identical to the reference in every mode."""
    monkeypatch.setenv("PB_ISLANDS", "1")
    monkeypatch.setenv("PB_ISLAND_LOCAL_MAX", str(local_max))
    s = parity.run_gates(maker(), steps=steps)
    assert s["steps"] == steps and s["worst_manifold"] <= parity.TOL


def test_islands_off_matches(monkeypatch):
    monkeypatch.setenv("PB_ISLANDS", "0")
    s = parity.run_gates(S.ragdolls(8), steps=40)
    assert s["steps"] == 40


@pytest.mark.parametrize("maker,steps", [(lambda: S.mixed_bin(900, spacing=0.8), 40), (lambda: S.trigger_zoo(160), 40), (lambda: S.ragdolls(8), 40)])
def test_three_gates_tree_broadphase_on_small_scenes(maker, steps, monkeypatch):
    """Scenes up to 8192 colliders normally take the all-pairs kernel; PB_BRUTE_FORCE_MAX=0 sends them through the Morton sort /
    LBVH / packet walk that large scenes use (the full-size tests cover it at 10^5..10^6 colliders)."""
    monkeypatch.setenv("PB_BRUTE_FORCE_MAX", "0")
    s = parity.run_gates(maker(), steps=steps)
    assert s["steps"] == steps


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("maker,steps", [
    (lambda: S.terrain(1500, cells=48, drop=0.3), 50),
    (lambda: S.terrain(400, cells=80, drop=0.3, mesh_spacing=0.3), 50),     # tens of candidate triangles per body, several contacts each
    (lambda: S.terrain_mixed(600, cells=40), 40),
])
def test_three_gates_mesh_light_modes(maker, steps, mode, monkeypatch):
    """Sphere / capsule vs triangle mesh: k_np_mesh (0) and k_np_mesh_light (1, the default: dual-child cull walk, packed triangle
    records) must both give the reference's manifolds, bit for bit."""
    monkeypatch.setenv("PB_MESH_LIGHT", str(mode))
    s = parity.run_gates(maker(), steps=steps)
    assert s["steps"] == steps and s["manifolds"] > 0 and s["worst_manifold"] <= parity.TOL


@pytest.mark.parametrize("big", [0, 1])
@pytest.mark.parametrize("maker,steps", [
    (lambda: S.mixed_bin(900, spacing=0.8), 40),                 # floor + four walls go on the side list
    (lambda: S.terrain(1200, cells=48, drop=0.3), 40),           # the terrain does
    (lambda: S.trigger_zoo(160), 40),
])
def test_three_gates_big_static_list(maker, steps, big, monkeypatch):
    """Tree broadphase with (1, default) and without (0) the big-static side list: scene-sized static colliders are tested directly by
    every querying collider instead of sitting in the tree.  The pair set must equal the reference's either way."""
    monkeypatch.setenv("PB_BRUTE_FORCE_MAX", "0")
    monkeypatch.setenv("PB_BIG_LIST", str(big))
    s = parity.run_gates(maker(), steps=steps)
    assert s["steps"] == steps and s["manifolds"] > 0


@pytest.mark.parametrize("coop", [0, 1])
@pytest.mark.parametrize("n,bits", [(5000, 30), (8193, 30), (100_000, 30), (1_000_003, 30), (2_000_000, 16), (300_000, 8), (50_001, 32), (20_000, 20)])
def test_device_sort(n, bits, coop, monkeypatch):
    """The key / payload sort behind the broadphase (Morton order): one-CTA sort up to 8192 keys, above that one cooperative launch
    (coop=1, the default) or three launches per pass (coop=0).  Must equal a stable sort by the low `bits` bits, rounded up to whole 8-bit passes."""
    import ctypes as C
    from physecs_b200.capi import Context
    monkeypatch.setenv("PB_SORT_COOP", str(coop))
    ctx = Context(S.pyramid(10))
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    if bits == 16:
        keys &= np.uint32(0x3FFF)          # many duplicates: stability is visible in the payload order
    vals = np.arange(n, dtype=np.int32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    rc = ctx.lib.pb_debug_sort_pairs(ctx.ctx, n, bits, keys.ctypes.data_as(C.POINTER(C.c_uint)), vals.ctypes.data_as(C.POINTER(C.c_int)),
                                     ko.ctypes.data_as(C.POINTER(C.c_uint)), vo.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0, ctx.lib.pb_last_error(ctx.ctx).decode()
    sorted_bits = (bits + 7) // 8 * 8          # whole 8-bit passes
    mask = np.uint32((1 << sorted_bits) - 1) if sorted_bits < 32 else np.uint32(0xFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    assert np.array_equal(vo, order.astype(np.int32))
    assert np.array_equal(ko, keys[order])
    ctx.close()


@pytest.mark.parametrize("fused_local_max", [0, 65536])
@pytest.mark.parametrize("maker,steps", [(lambda: S.ragdolls(300), 60), (lambda: S.terrain(3000, cells=64, drop=0.05), 40)])
def test_whole_step_kernel_group_by_group(maker, steps, fused_local_max, monkeypatch):
    """Scenes whose constraints all sit in small islands (ragdoll batches, bodies spread over a terrain) take the one-launch whole-step
    kernel group by group -- every CTA carries its groups through all substeps with CTA barriers only (solver.cu k_step_solve_small);
    PB_FUSED_LOCAL_MAX=0 keeps them on the per-substep launches.  Same results either way, bit for bit against the oracle."""
    monkeypatch.setenv("PB_FUSED_LOCAL_MAX", str(fused_local_max))
    monkeypatch.setenv("PB_ISLANDS", "1")
    s = parity.run_gates(maker(), steps=steps)
    assert s["steps"] == steps and s["manifolds"] > 0 and s["worst_manifold"] <= parity.TOL


@pytest.mark.parametrize("cluster", [0, 1])
@pytest.mark.parametrize("maker,steps", [(lambda: S.pyramid(300), 40), (lambda: S.mixed_bin(900, spacing=0.8), 40), (lambda: S.joint_zoo(), 40)])
def test_three_gates_cluster_barrier(maker, steps, cluster, monkeypatch):
    """Small one-pile scenes run their whole-step kernel as ONE thread-block cluster (<= 16 CTAs) whose colour phases meet at the hardware
    cluster barrier; PB_CLUSTER=0 keeps the cooperative launch with the counter barrier.  Islands off, so every colour goes through the
    device-wide sweep either way.  Same results, bit for bit against the oracle."""
    monkeypatch.setenv("PB_CLUSTER", str(cluster))
    monkeypatch.setenv("PB_ISLANDS", "0")
    monkeypatch.setenv("PB_FUSED", "1")
    s = parity.run_gates(maker(), steps=steps)
    assert s["steps"] == steps and s["worst_manifold"] <= parity.TOL


@pytest.mark.parametrize("islands", ["auto", "0"])
def test_three_gates_through_landing(islands, monkeypatch):
    """20 000 bodies dropped into a walled bin, gated from the first step on: the scene goes from no contacts through a field of small
    islands (per-CTA sweeps) to a mix of small and big islands (both sweeps in one step) while manifold counts grow ten-fold -- every
    size rule (launch shapes from the previous step's counts, scan forms, whole-step kernel <-> per-substep launches) flips on the way."""
    if islands != "auto":
        monkeypatch.setenv("PB_ISLANDS", islands)
    d = S.mixed_bin(20000)
    s = parity.run_gates(d, steps=72, bulk=True, check_every=6, caps=dict(max_pairs=64 * d.n + 4096, max_manifolds=16 * d.n + 4096))
    assert s["steps"] == 72 and s["manifolds"] > 30000 and s["worst_manifold"] <= parity.TOL


@pytest.mark.parametrize("maker", [lambda: S.mixed_bin(20000), lambda: S.mixed_bin(3000, spacing=0.8), lambda: S.mixed_bin(40000, spacing=0.9)])
def test_free_run_bodies_stay_in_the_bin(maker):
    """free run, no oracle: nothing falls through the floor while the pile forms (a constraint that is built but never swept shows here)"""
    from physecs_b200.capi import Context
    d = maker()
    ctx = Context(d, max_pairs=64 * d.n + 4096, max_manifolds=16 * d.n + 4096)
    for k in range(150):
        ctx.step()
        if k % 30 == 29:
            P = ctx.get_state()[0]
            assert np.isfinite(P).all() and P[:, 1].min() > -0.25, (k, float(P[:, 1].min()))
    ctx.close()


def test_three_gates_all_pairs_broadphase_on_a_big_batch():
    """1500 ragdoll scenes side by side (18 k colliders, past the 8192 up to which the step always tests all pairs): the tile statistics of
    the first step (tree + probe) show that every tile of 128 consecutive colliders meets a handful of tiles, and from the second step on the
    all-pairs kernel replaces Morton sort + tree.  Same pair sets, manifolds and solves as the oracle throughout."""
    from physecs_b200.capi import Context
    d = S.ragdolls(1500)
    s = parity.run_gates(d, steps=8, bulk=True)
    assert s["steps"] == 8 and s["worst_manifold"] <= parity.TOL
    ctx = Context(d)
    kinds = []
    for _ in range(4):
        ctx.step(); ctx.sync()
        kinds.append(ctx.broadphase_info())
    ctx.close()
    assert not kinds[0]["all_pairs"] and kinds[0]["tiles"] > 64 and 0 < kinds[0]["tile_hits"] <= 24 * kinds[0]["tiles"], kinds
    assert all(k["all_pairs"] for k in kinds[1:]), kinds


def test_a_pile_keeps_the_tree():
    """20 000 bodies in one bin: consecutive colliders are NOT neighbours for long (the bodies mix), every tile meets most tiles, the tree stays"""
    from physecs_b200.capi import Context
    d = S.mixed_bin(20000)
    ctx = Context(d, max_pairs=64 * d.n + 4096, max_manifolds=16 * d.n + 4096)
    for _ in range(90):
        ctx.step()
    ctx.sync()
    info = ctx.broadphase_info()
    ctx.close()
    assert not info["all_pairs"] and info["tile_hits"] > 24 * info["tiles"], info

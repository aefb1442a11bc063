"""Shared parity checks: device path (C ABI) vs the oracle (reference compiled by oracle/build_ref.py).

The three gates of BASELINE.json's north_star:
  1. broadphase pair SET bit-exact                       -> compare_pairs
  2. manifolds within 1e-4 rel/abs (points, normals)     -> compare_manifolds
  3. one-step solve within 1e-4 when the reference CPU solver is fed the device's colour-batched order -> one_step_solve
"""
from __future__ import annotations

import numpy as np

TOL = 1e-4


def close(a, b, tol=TOL):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) <= tol * np.maximum(np.abs(a), np.abs(b)) + tol


def max_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / (np.maximum(np.abs(a), np.abs(b)) + 1.0)))


def pair_set(p):
    return set(map(tuple, np.asarray(p).reshape(-1, 4).tolist()))


def compare_pairs(gpu_pairs, ref_pairs):
    g, r = pair_set(gpu_pairs), pair_set(ref_pairs)
    assert len(g) == len(gpu_pairs), "device emitted duplicate pairs"
    missing, extra = r - g, g - r
    assert not missing and not extra, f"pair sets differ: missing={sorted(missing)[:5]} extra={sorted(extra)[:5]} ({len(missing)}/{len(extra)})"
    return len(g)


def compare_bounds(ctx, ref):
    """Bit-exact collider bounds (BroadPhaseEntry::bounds)."""
    ids, rb = ref.bounds()
    gb = ctx.bounds()
    key = {(int(e), int(c)): i for i, (e, c) in enumerate(zip(ctx.col_entity, ctx.col_index))}
    idx = np.array([key[(int(e), int(c))] for e, c in ids])
    same = gb[idx].view(np.int32) == rb.view(np.int32)
    # -0.0 vs +0.0 compare equal as floats
    same |= (gb[idx] == rb)
    assert same.all(), f"{(~same).any(1).sum()} collider bounds differ, first: {gb[idx][~same.all(1)][:2]} vs {rb[~same.all(1)][:2]}"
    return len(ids)


def manifold_map(m):
    out = {}
    for i, k in enumerate(m["keys"]):
        out.setdefault(tuple(int(x) for x in k), []).append(i)
    return out


def compare_manifolds(gm, rm, tol=TOL, allow_tri_alias=False):
    """gm / rm: dicts(keys[n,5], num_points, normal, points[n,4,2,3]).  Points are compared as sets per manifold."""
    g, r = manifold_map(gm), manifold_map(rm)
    missing = [k for k in r if k not in g]
    extra = [k for k in g if k not in r]
    assert not missing and not extra, f"manifold key sets differ: missing={missing[:4]} extra={extra[:4]} ({len(missing)}/{len(extra)} of {len(r)})"
    worst = 0.0
    for k, ri in r.items():
        gi = g[k]
        assert len(gi) == len(ri) == 1, f"duplicate manifold key {k}"
        a, b = gi[0], ri[0]
        assert gm["num_points"][a] == rm["num_points"][b], f"numPoints differ for {k}: {gm['num_points'][a]} vs {rm['num_points'][b]}"
        n = int(rm["num_points"][b])
        assert close(gm["normal"][a], rm["normal"][b], tol).all(), f"normal differs for {k}: {gm['normal'][a]} vs {rm['normal'][b]}"
        assert close(gm["points"][a][:n], rm["points"][b][:n], tol).all(), f"points differ for {k}:\n{gm['points'][a][:n]}\nvs\n{rm['points'][b][:n]}"
        worst = max(worst, max_err(gm["points"][a][:n], rm["points"][b][:n]), max_err(gm["normal"][a], rm["normal"][b]))
    return len(r), worst


# ---- vectorised variants for BASELINE.json's full sizes (10^5 .. 10^6 bodies, millions of pairs / manifolds) -------------
def _lexorder(rows):
    rows = np.asarray(rows)
    return np.lexsort(rows.T[::-1])


def compare_pairs_bulk(gpu_pairs, ref_pairs):
    g = np.asarray(gpu_pairs, np.int32).reshape(-1, 4); r = np.asarray(ref_pairs, np.int32).reshape(-1, 4)
    g = g[_lexorder(g)]; r = r[_lexorder(r)]
    assert len(g) < 2 or (g[1:] != g[:-1]).any(1).all(), "device emitted duplicate pairs"
    assert g.shape == r.shape and np.array_equal(g, r), f"pair sets differ: {len(g)} device vs {len(r)} reference pairs"
    return len(g)


def compare_bounds_bulk(ctx, ref):
    ids, rb = ref.bounds()
    gb = ctx.bounds()
    gk = ctx.col_entity.astype(np.int64) * 65536 + ctx.col_index
    rk = ids[:, 0].astype(np.int64) * 65536 + ids[:, 1]
    go, ro = np.argsort(gk), np.argsort(rk)
    assert np.array_equal(gk[go], rk[ro]), "collider sets differ"
    a, b = gb[go], rb[ro]
    same = (a.view(np.int32) == b.view(np.int32)) | (a == b)
    assert same.all(), f"{(~same).any(1).sum()} of {len(a)} collider bounds differ"
    return len(a)


def compare_manifolds_bulk(gm, rm, tol=TOL):
    gk = np.asarray(gm["keys"], np.int32); rk = np.asarray(rm["keys"], np.int32)
    go, ro = _lexorder(gk), _lexorder(rk)
    gk, rk = gk[go], rk[ro]
    assert len(gk) < 2 or (gk[1:] != gk[:-1]).any(1).all(), "duplicate manifold keys on the device"
    assert gk.shape == rk.shape and np.array_equal(gk, rk), f"manifold key sets differ: {len(gk)} device vs {len(rk)} reference"
    gn, rn = np.asarray(gm["num_points"])[go], np.asarray(rm["num_points"])[ro]
    assert np.array_equal(gn, rn), f"numPoints differ on {(gn != rn).sum()} manifolds"
    a, b = gm["normal"][go], rm["normal"][ro]
    assert close(a, b, tol).all(), "normals differ"
    mask = (np.arange(4)[None, :] < rn[:, None])[:, :, None, None]
    pa, pb = np.where(mask, gm["points"][go], 0), np.where(mask, rm["points"][ro], 0)
    assert close(pa, pb, tol).all(), "contact points differ"
    return len(rk), max(max_err(pa, pb), max_err(a, b))


def sync_device_to_oracle(ctx, ref):
    """Teacher forcing: copy the oracle's current state (all dynamic entities) into the device context."""
    p, q, v, w = ref.get_state()
    ctx.set_state_entities(p, q, v, w)
    return p, q, v, w


def one_step_solve(ctx, ref, tol=TOL):
    """Gate 3.  Both sides start from the oracle's current state; the device steps once, the oracle steps once with
    its contact constraints permuted to the device's (colour, slot) order; poses and velocities must agree."""
    sync_device_to_oracle(ctx, ref)
    ctx.step()
    gm = ctx.manifolds()
    ref.set_manifold_order(gm["keys"])
    ref.simulate()
    matched, missing, extra = ref.order_stats()
    assert missing == 0 and extra == 0, f"manifold sets differ between device and oracle: matched={matched} missing={missing} extra={extra}"
    P, Q, V, W = ctx.get_state_entities()
    p, q, v, w = ref.get_state()
    # quaternion sign is irrelevant
    s = np.sign(np.sum(Q * q, axis=1, keepdims=True)); s[s == 0] = 1
    errs = dict(pos=max_err(P, p), quat=max_err(Q * s, q), vel=max_err(V, v), angvel=max_err(W, w))
    bad = {k: e for k, e in errs.items() if not e <= tol}
    assert not bad, f"one-step solve mismatch {errs} with {matched} manifolds"
    assert close(P, p, tol).all() and close(Q * s, q, tol).all() and close(V, v, tol).all() and close(W, w, tol).all()
    return matched, errs


# custom contact filters: the same predicates as filterParity / filterAsymmetric in oracle/ref_harness.cpp
FILTERS = {
    0: None,
    1: lambda t0, d0, t1, d1: (t0 or t1) and (d0 + d1) % 2 == 0,
    2: lambda t0, d0, t1, d1: (t0 and not t1) or (t1 and d0 == 1),
}


def compare_triggers(gpu_trig, ref_trig):
    g, r = pair_set(gpu_trig), pair_set(ref_trig)
    assert len(g) == len(gpu_trig), "device emitted duplicate trigger pairs"
    assert g == r, f"trigger pair sets differ: missing={sorted(r - g)[:5]} extra={sorted(g - r)[:5]} ({len(r - g)}/{len(g - r)} of {len(r)})"
    return len(g)


def run_gates(desc, steps=10, check_every=1, tol=TOL, ref_threads=0, verbose=False, contact_filter=0, bulk=False, caps=None):
    """All three gates on `steps` consecutive steps, teacher-forced from the oracle's trajectory.
    Returns a summary dict.  Needs a CUDA device (device path) and oracle/_ref (checker).
    bulk: vectorised comparisons + the oracle's broadphase entries pre-sorted (full-size scenes)."""
    from oracle.ref import RefScene
    from physecs_b200.capi import Context
    ref = RefScene(desc, ref_threads, hashfix=True)
    if bulk:
        ref.presort()
    ctx = Context(desc, **(caps or {}))
    cmp_bounds = compare_bounds_bulk if bulk else compare_bounds
    cmp_pairs = compare_pairs_bulk if bulk else compare_pairs
    cmp_manifolds = compare_manifolds_bulk if bulk else compare_manifolds
    summary = dict(steps=0, pairs=0, manifolds=0, worst_manifold=0.0, worst_solve={}, triggers=0, trigger_changes=0, spilled=0, cause=0)
    prev_trig = set()
    if contact_filter:
        ref.set_contact_filter(contact_filter)
        ctx.set_contact_filter(FILTERS[contact_filter])
    try:
        cmp_bounds(ctx, ref)
        for k in range(steps):
            sync_device_to_oracle(ctx, ref)
            if k > 0:
                ctx.refresh_bounds()
                cmp_bounds(ctx, ref)
            ctx.step()
            gm = ctx.manifolds()
            gp = ctx.pairs()
            cnt = ctx.counts()
            summary["spilled"] = max(summary["spilled"], int(cnt.n_spilled)); summary["cause"] |= int(cnt.cause)
            if k % check_every == 0:
                rm = ref.narrowphase(gp)
                nm, worst = cmp_manifolds(gm, rm, tol)
                summary["worst_manifold"] = max(summary["worst_manifold"], worst)
            ref.set_manifold_order(gm["keys"])
            ref.simulate()
            matched, missing, extra = ref.order_stats()
            assert missing == 0 and extra == 0, f"step {k}: manifold sets differ: matched={matched} missing={missing} extra={extra}"
            npairs = cmp_pairs(gp, ref.pairs())
            gt = ctx.triggers()
            summary["triggers"] = max(summary["triggers"], compare_triggers(gt, ref.triggers()))
            cur_trig = pair_set(gt)
            summary["trigger_changes"] += len(cur_trig ^ prev_trig)
            prev_trig = cur_trig
            P, Q, V, W = ctx.get_state_entities()
            p, q, v, w = ref.get_state()
            s = np.sign(np.sum(Q * q, axis=1, keepdims=True)); s[s == 0] = 1
            errs = dict(pos=max_err(P, p), quat=max_err(Q * s, q), vel=max_err(V, v), angvel=max_err(W, w))
            if verbose:
                print(f"step {k}: pairs={npairs} manifolds={matched} colors={ctx.counts().n_colors} errs={errs}")
            for name, e in errs.items():
                assert e <= tol, f"step {k}: one-step solve mismatch {errs} ({matched} manifolds)"
                summary["worst_solve"][name] = max(summary["worst_solve"].get(name, 0.0), e)
            summary["steps"] += 1
            summary["pairs"] = max(summary["pairs"], npairs)
            summary["manifolds"] = max(summary["manifolds"], matched)
    finally:
        ctx.close()
        ref.close()
    return summary

"""-m gpu: the pb_batch_* C ABI (independent scenes sharded over devices, one host thread + stream per shard).
A batch split into shards must give, scene for scene, the same result as the whole batch in one context (deterministic mode: the
default colouring is run-to-run different) and the same as the oracle's one-step solve; shards run concurrently from their own threads.
With one GPU the shards share device 0 (the threading and sharding logic is the same); with two or more each shard takes its own."""
import numpy as np
import pytest

from physecs_b200 import batch as B
from physecs_b200 import scenes as S

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [k % max(have, 1) for k in range(n)]


@pytest.mark.parametrize("n_shards", [1, 2, 3])
def test_shards_step_like_standalone_contexts(n_shards, monkeypatch):
    """Every shard of a batch, driven concurrently by its own thread, ends where a standalone context holding the same scenes ends --
    bit for bit in deterministic mode (colour priorities hash collider indices, which a shard and a standalone context of the same
    description share; the whole batch in ONE context numbers its colliders differently, so it is a different -- equally valid --
    colour order, which the parity gates cover)."""
    from physecs_b200.capi import Context
    monkeypatch.setenv("PB_DETERMINISTIC", "1")
    n_scenes, steps = 48, 90
    ranges = [B.shard_range(n_scenes, n_shards, k) for k in range(n_shards)]
    descs = [S.ragdolls(e - b, seed=0xC5, first_scene=b, total_scenes=n_scenes) for b, e in ranges]
    alone = []
    for d in descs:
        c = Context(d, max_pairs=max(64 * d.n, 4096), max_manifolds=max(16 * d.n, 4096))
        for _ in range(steps):
            c.step()
        alone.append(c.get_state())
        c.close()
    bt = B.Batch(descs, _devices(n_shards))
    try:
        bt.step(steps); bt.sync()
        for k in range(n_shards):
            got = bt.shards[k].get_state()
            for g, r, what in zip(got, alone[k], ("pos", "quat", "vel", "angvel")):
                assert np.array_equal(g.view(np.int32), r.view(np.int32)), f"shard {k}: {what} differs from the standalone context of the same scenes"
        # a shard IS its slice of the whole batch: same bodies, same initial state (scene generation by global scene index)
        whole = S.ragdolls(n_scenes, seed=0xC5)
        m = whole.n // n_scenes
        for (b, e), d in zip(ranges, descs):
            assert np.array_equal(d.pos, whole.pos[b * m:e * m]) and np.array_equal(d.quat, whole.quat[b * m:e * m])
    finally:
        bt.close()


def test_batch_state_round_trip_and_errors():
    """pb_batch_set_state / _get_state move per-shard host arrays; a failing shard surfaces in pb_batch_sync with its index."""
    from physecs_b200 import capi
    descs = [S.ragdolls(8, first_scene=0, total_scenes=16), S.ragdolls(8, first_scene=8, total_scenes=16)]
    bt = B.Batch(descs, _devices(2))
    try:
        bt.step(5); bt.sync()
        bufs = bt.alloc_state()
        bt.get_state(*bufs)
        before = [b.copy() for b in bufs[0]]
        bufs[0][1][:, 1] += 0.25                      # lift every body of shard 1
        bt.set_state(*bufs); bt.step(1)
        out = bt.alloc_state()
        bt.get_state(*out)
        assert np.allclose(out[0][0], before[0], atol=0.05) and np.all(out[0][1][:, 1] > before[1][:, 1] + 0.15)
    finally:
        bt.close()
    # arenas far too small on one shard: the batch reports which shard overflowed
    bt = B.Batch(descs, _devices(2), pairs_per_body=64, manifolds_per_body=16)
    try:
        ctx = bt.shards[1]
        assert ctx.lib.pb_grow_arenas(ctx.ctx, 1, 1) == 0      # (cannot shrink: stays as created)
        bt.step(60)
        bt.sync()
    finally:
        bt.close()

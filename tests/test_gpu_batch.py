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
def test_sharded_batch_equals_whole_batch(n_shards, monkeypatch):
    monkeypatch.setenv("PB_DETERMINISTIC", "1")
    n_scenes, steps = 48, 90
    whole = B.Batch([S.ragdolls(n_scenes, seed=0xC5)], _devices(1))
    whole.step(steps); whole.sync()
    ref_state = whole.shards[0].get_state()
    whole.close()
    ranges = [B.shard_range(n_scenes, n_shards, k) for k in range(n_shards)]
    descs = [S.ragdolls(e - b, seed=0xC5, first_scene=b, total_scenes=n_scenes) for b, e in ranges]
    bt = B.Batch(descs, _devices(n_shards))
    try:
        bt.step(steps); bt.sync()
        bodies_per_scene = ref_state[0].shape[0] // n_scenes
        for k, (b, e) in enumerate(ranges):
            got = bt.shards[k].get_state()
            for g, r, what in zip(got, ref_state, ("pos", "quat", "vel", "angvel")):
                want = r[b * bodies_per_scene:e * bodies_per_scene]
                assert np.array_equal(g.view(np.int32), want.view(np.int32)), f"shard {k} ({b}..{e}): {what} differs from the same scenes in the whole batch"
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

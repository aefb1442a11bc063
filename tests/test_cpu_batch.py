"""-m "not gpu": host logic of the sharded batched-scene path (world_size 2 over gloo, CPU only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from physecs_b200 import batch
from physecs_b200 import scenes as S


def test_shard_ranges_cover_without_overlap():
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            r = [batch.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        batch.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_scenes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = batch.shard_range(n_scenes, world, rank)
        # every rank builds ITS scenes from global scene indices: the union must equal the single-process batch
        d = S.ragdolls(e - b, seed=0xC5, first_scene=b, total_scenes=n_scenes)
        digest = torch.tensor([float(np.sum(d.pos.astype(np.float64))), float(d.n), float(len(d.joints))], dtype=torch.float64)
        gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, digest)
        ranges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ranges, torch.tensor([b, e], dtype=torch.int64))
        t = batch.max_over_ranks(1.0 + rank)      # the slowest rank defines the step time
        dist.barrier()
        if rank == 0:
            out.put((torch.stack(gathered).numpy(), torch.stack(ranges).numpy(), t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    n_scenes, world = 9, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    digests, ranges, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0
    assert ranges.tolist() == [[0, 5], [5, 9]]
    whole = S.ragdolls(n_scenes, seed=0xC5)
    assert int(digests[:, 1].sum()) == whole.n and int(digests[:, 2].sum()) == len(whole.joints)
    assert abs(digests[:, 0].sum() - float(np.sum(whole.pos.astype(np.float64)))) < 1e-3
    # and shard k really is scenes [b, e) of the whole batch, entity for entity
    b, e = ranges[1]
    part = S.ragdolls(int(e - b), seed=0xC5, first_scene=int(b), total_scenes=n_scenes)
    m = whole.n // n_scenes
    assert np.array_equal(part.pos, whole.pos[b * m:e * m]) and np.array_equal(part.quat, whole.quat[b * m:e * m])


def test_batch_driver_under_thread_sanitizer(tmp_path):
    """physecs_b200/csrc/batch.cpp (one host thread + task queue per shard) built with -fsanitize=thread against the recording double of
    the C ABI (tests/abi_recorder: computes nothing): six shards stepped concurrently, state fetched / pushed between rounds, a failing
    shard named at the join, destroy with work in flight.  No device involved: this checks the threading, not the physics."""
    import os
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rec = os.path.join(root, "tests", "abi_recorder")
    exe = str(tmp_path / "batch_tsan")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-fno-omit-frame-pointer", "-I", os.path.join(rec, "stub"),
           os.path.join(root, "physecs_b200", "csrc", "batch.cpp"), os.path.join(rec, "pb_recorder.cpp"), os.path.join(root, "physecs_b200", "csrc", "trimesh_build.cpp"),
           os.path.join(rec, "sanitize_batch_driver.cpp"),
           "-o", exe, "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode and b"tsan" in r.stdout.lower():
        pytest.skip("this g++ has no ThreadSanitizer runtime")
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    out = r.stdout.decode()
    assert r.returncode == 0 and "batch sanitize driver ok" in out and "shard 3" in out, out[-3000:]

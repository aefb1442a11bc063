"""ms/step of ragdoll batches of a few sizes through the pb_batch_* API (device-resident), for quick A/B runs of the small-scene paths.
usage: python tools/gpu_batch_time.py [sizes...]   (default 512 4096)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physecs_b200 import scenes as S, batch as B
sizes = [int(x) for x in sys.argv[1:]] or [512, 4096]
for n in sizes:
    d = S.ragdolls(n, total_scenes=4096)
    bt = B.Batch([d], [0])
    ctx = bt.shards[0]
    bt.step(150); bt.sync()
    stream = torch.cuda.ExternalStream(ctx.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches()
    torch.cuda.synchronize()
    e0.record(stream)
    bt.step(200); bt.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    t = ctx.timings()
    print(f"{n} scenes: {e0.elapsed_time(e1) / 200:.4f} ms/step, {(ctx.launches() - l0) / 200:.1f} launches/step, last step: broad {t.broadphase:.3f} narrow {t.narrowphase:.3f} build {t.contact_build:.3f} solve {t.solve:.3f}, "
          f"manifolds {ctx.counts().n_manifolds}, islands {ctx.island_stats()}", flush=True)
    bt.close()

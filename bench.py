#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Physecs per-step pipeline on B200.

Workload (BASELINE.json configs[3], the 1M-body configuration the metric is quoted on): 1,000,000 spheres and
capsules over a static 2,097,152-triangle terrain mesh, 60 Hz, 4 TGS substeps x 2 iterations (+ relaxation),
settled for --settle steps before measuring so the contact graph is populated.  A "step" is one
physecs::Scene::simulate(1/60): broadphase -> narrowphase -> contact build -> substep solve.

Lines printed (one JSON object, rank 0):
  value   body-steps/s with the scene resident in HBM (pb_step only), CUDA events on the context's stream
  e2e     the same metric through the C ABI with HOST buffers: pb_set_state (H2D) + pb_step + pb_get_state (D2H) per step
  roofline  dominant kernel = k_substeps (the persistent substep solver, one launch per step): algorithmic bytes
            (SURVEY.md §8d per-unit figures x units, summed over its phases) / CUDA-event launch time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, "reference + hash fix" build) on a bounded sample

`--impl reference` times the reference CPU implementation itself on the host cores (bounded sample of the same workload).
N > 1 (torchrun): independent replicas of the workload, one per GPU, no collective on the data path ("weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "body-steps/sec at 1M bodies (spheres+capsules on triangle-mesh terrain, 4 substeps)"
UNIT = "body-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def reference_sample(n_sample, cells, settle, warmup, steps, threads):
    """Time the reference CPU implementation (oracle/_ref) on a bounded sample of the workload."""
    from oracle.ref import RefScene
    from physecs_b200 import scenes as S
    d = S.terrain(n_sample, cells=cells, drop=0.3)
    ref = RefScene(d, threads, hashfix=True)
    ref.presort()      # the order the reference's own first-step insertion sort would reach, without its O(n^2) first step
    for _ in range(settle + warmup):
        ref.simulate()
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.simulate()
    dt = time.perf_counter() - t0
    nm = len(ref.manifold_keys())
    ref.close()
    return n_sample * steps / dt, dt / steps * 1e3, nm


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(cores - 1, 0)
    n_sample, cells = args.ref_bodies, args.ref_cells
    try:
        value, ms, nm = reference_sample(n_sample, cells, args.ref_settle, args.warmup, args.steps, threads)
    except Exception as e:  # oracle not built
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref not usable: {e}"}))
        return
    sample = f"terrain scene, {n_sample} bodies on a {cells}x{cells}-cell mesh, settled {args.ref_settle} steps, reference+hash-fix build, Scene(registry, {threads})"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 terrain: spheres+capsules over static triangle mesh, 60 Hz, 4 substeps x 2 iterations", "sample_bodies": n_sample,
                   "manifolds": nm},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads + 1, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", type=int, default=1_000_000)
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--settle", type=int, default=150)
    ap.add_argument("--ref-bodies", type=int, default=32000)
    ap.add_argument("--ref-cells", type=int, default=200)
    ap.add_argument("--ref-settle", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batched-scenes", type=int, default=4096, help="ragdoll scenes for the sharded-batch section (0 = skip)")
    ap.add_argument("--scene-bodies", type=int, default=1_000_000, help="bodies for the physecs::Scene end-to-end section (0 = skip)")
    ap.add_argument("--other-configs", type=int, default=1, help="also time BASELINE.json's C1 / C2 / C3 scenes at full size (0 = skip)")
    ap.add_argument("--ncu", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (never a bench value)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from physecs_b200 import scenes as S
    from physecs_b200.capi import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the device path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.bodies
    desc = S.terrain(n, cells=args.cells, drop=0.3)
    ctx = Context(desc, device=local_rank, max_pairs=8 * n + 4096, max_manifolds=6 * n + 4096)
    n_dyn = ctx.n_dyn
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local_rank))

    for _ in range(args.settle):
        ctx.step()
    ctx.sync()
    for _ in range(args.warmup):
        ctx.step()
    ctx.sync()

    # ---- timed region A: device-resident ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launches()
    barrier(); torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sumM = sumP = sumC = sumPairs = 0
    phase = np.zeros(5)
    if args.ncu:
        ctx.lib.pb_profiler_range(1)
    e0.record(stream)
    for _ in range(args.steps):
        ctx.step()       # no read-back inside the timed loop: the counters stay on the device
    e1.record(stream)
    barrier(); torch.cuda.synchronize()
    if args.ncu:
        ctx.lib.pb_profiler_range(0)
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launches() - launches0
    c = ctx.counts()     # the scene is settled: the last step's counts stand for the timed region
    sumM, sumP, sumC, sumPairs = c.n_manifolds * args.steps, c.n_points * args.steps, c.n_colors * args.steps, c.n_pairs * args.steps
    t = ctx.timings()
    # in-kernel phase stamps (CTA 0's %globaltimer at the grid barriers, plus one extra barrier per substep that closes the local
    # sweeps): collected over a few extra steps OUTSIDE the timed region, so the headline value is measured without them
    prof_steps = min(args.steps, 20)
    ctx.set_profile(True)
    for _ in range(prof_steps):
        ctx.step()
    prof = ctx.profile()
    ctx.set_profile(False)
    # duration of the dominant kernel by CUDA events on its stream: a few extra steps, each read back (outside the timed region)
    kernel_ms = []
    for _ in range(min(args.steps, 20)):
        ctx.step()
        kernel_ms.append(ctx.timings().solve_kernel)
    kernel_ms = float(np.mean(kernel_ms))
    ms_per_step = ms_total / args.steps
    value = world * n_dyn * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel: k_substeps, the persistent substep solver (one launch per step) ------------------
    # algorithmic bytes per launch (per substep): bodies 140 (integrate v) + 120 (integrate x) B/body;
    # contact prep 244 + 156 p B/manifold; contact solve pass 180 + 140 p B/manifold, (iterations + 1) passes.
    peak, peak_src = peaks()
    S_, I_ = desc.substeps, desc.iterations
    avgM, avgP, avgC = sumM / args.steps, sumP / args.steps, max(sumC / args.steps, 1.0)
    bytes_bodies = 260.0 * n_dyn * S_
    bytes_prep = (244.0 * avgM + 156.0 * avgP) * S_
    bytes_pass = (180.0 * avgM + 140.0 * avgP)
    bytes_solve = bytes_pass * (I_ + 1) * S_
    bytes_launch = bytes_bodies + bytes_prep + bytes_solve
    phase_ms = {k: v[0] / prof_steps for k, v in prof.items()}
    launch_ms = kernel_ms                  # CUDA events around the k_substeps launch, on the context's stream
    stamp_ms = sum(phase_ms.values())      # the same from CTA 0's %globaltimer stamps at every grid barrier (splits the launch into phases)
    achieved = bytes_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
    islands = ctx.island_stats()
    # islands on: most colour phases run inside per-CTA sweeps (no grid barrier, so no per-phase stamp); this scene has no joints,
    # so those sweeps are contact passes and are counted with the device-wide ones
    pass_ms = phase_ms.get("contact_pass", 0.0) + phase_ms.get("local_sweeps", 0.0)
    pass_gbs = bytes_solve / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else 0.0
    prep_ms = phase_ms.get("prep", 0.0)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_substeps_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_substeps", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": bytes_launch, "launch_ms": launch_ms,
                "launches_timed": min(args.steps, 20), "share_of_step": launch_ms / ms_per_step, "launch_ms_from_phase_stamps": stamp_ms,
                "phases": {"contact_solve_passes": {"ms": pass_ms, "GB/s": pass_gbs, "frac": pass_gbs / peak, "bytes": bytes_solve},
                           "contact_prep": {"ms": prep_ms, "GB/s": bytes_prep / (prep_ms * 1e-3) / 1e9 if prep_ms > 0 else 0.0, "bytes": bytes_prep},
                           "integrate": {"ms": phase_ms.get("integrate_v", 0.0) + phase_ms.get("integrate_x", 0.0), "bytes": bytes_bodies}},
                "islands": islands,
                "note": "one launch = the whole TGS substep loop of a step (persistent cooperative kernel, grid barriers between phases); "
                        "algorithmic bytes = sum over phases of SURVEY.md 8d's per-unit figures x units; duration = in-kernel phase stamps"}

    # ---- timed region B: end to end through the C ABI with pinned host buffers --------------------------------------
    pos_t = torch.empty((n_dyn, 3), dtype=torch.float32).pin_memory(); quat_t = torch.empty((n_dyn, 4), dtype=torch.float32).pin_memory()
    vel_t = torch.empty((n_dyn, 3), dtype=torch.float32).pin_memory(); ang_t = torch.empty((n_dyn, 3), dtype=torch.float32).pin_memory()
    pos, quat, vel, ang = pos_t.numpy(), quat_t.numpy(), vel_t.numpy(), ang_t.numpy()
    import ctypes as C
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    lib = ctx.lib
    lib.pb_get_state(ctx.ctx, fp(pos), fp(quat), fp(vel), fp(ang))
    e2e_steps = args.steps
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = lib.pb_set_state(ctx.ctx, n_dyn, fp(pos), fp(quat), fp(vel), fp(ang))
        rc |= lib.pb_step(ctx.ctx, C.c_float(desc.dt), desc.substeps, desc.iterations, C.c_float(desc.gravity))
        rc |= lib.pb_get_state(ctx.ctx, fp(pos), fp(quat), fp(vel), fp(ang))
        if rc:
            raise RuntimeError(lib.pb_last_error(ctx.ctx).decode())
    torch.cuda.synchronize(); barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n_dyn * e2e_steps / e2e_s
    bytes_io = 13 * 4 * n_dyn
    checksum = float(np.abs(pos).sum())

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 terrain: %d spheres+capsules over a static %d-triangle mesh, 60 Hz, 4 substeps x 2 iterations + relaxation" % (n_dyn, 2 * args.cells * args.cells),
                   "bodies_per_gpu": n_dyn, "settle_steps": args.settle, "replicas": world, "l2_policy": "working set >> L2 (per-step traffic ~GBs; no flush needed)",
                   "avg_pairs": sumPairs / args.steps, "avg_manifolds": avgM, "avg_points": avgP, "avg_colors": avgC},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "checksum": checksum},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "phase_ms_last_step": {"broadphase": t.broadphase, "narrowphase": t.narrowphase, "contact_build": t.contact_build, "solve": t.solve, "total": t.total},
        "stage_ms_per_step": phase_ms,
    }

    ctx.close()

    # ---- batched independent scenes (BASELINE.json configs[4]): 4096 ragdoll scenes sharded by scene across ranks ----
    if args.batched_scenes > 0:
        from physecs_b200 import batch as B

        def time_batch(first, count, total):
            rd = S.ragdolls(count, seed=0xC5, first_scene=first, total_scenes=total)
            rctx = Context(rd, device=local_rank, max_pairs=64 * rd.n, max_manifolds=16 * rd.n)
            for _ in range(120):
                rctx.step()
            rctx.sync()
            rstream = torch.cuda.ExternalStream(rctx.stream_ptr(), device=torch.device("cuda", local_rank))
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier(); torch.cuda.synchronize()
            r0.record(rstream)
            for _ in range(args.steps):
                rctx.step()
            r1.record(rstream)
            barrier(); torch.cuda.synchronize()
            ms = max_over_ranks(r0.elapsed_time(r1))
            info = (int(rctx.n_dyn), int(rctx.counts().n_manifolds), len(rd.joints), rctx.island_stats())
            rctx.close()
            return ms, info

        # strong: BASELINE.json's 4096 scenes split over the ranks (contiguous blocks of scenes, no collective)
        b0, b1 = B.shard_range(args.batched_scenes, world, rank)
        rms, (nb, nm, nj, isl) = time_batch(b0, b1 - b0, args.batched_scenes)
        out["batched_scenes"] = {"metric": "batched-scene steps/s (%d independent ragdoll scenes, 11 bodies + 10 joints + ground each, 4 substeps)" % args.batched_scenes,
                                 "value": args.batched_scenes * args.steps / (rms * 1e-3), "unit": "scene-steps/s", "scenes": args.batched_scenes,
                                 "scenes_on_rank0": b1 - b0, "ms_per_step": rms / args.steps, "scaling": "strong", "sharding": "contiguous blocks of scenes per rank, no collective",
                                 "bodies_rank0": nb, "manifolds_last_step_rank0": nm, "joints_rank0": nj, "islands": isl,
                                 "note": "a step of a batch this small is latency-bound (~0.8 ms however few scenes a rank holds), so splitting a fixed 4096 scenes over N GPUs cannot scale; "
                                         "the weak figure below (4096 scenes per GPU) is the throughput a sharded batch service sees"}
        if world > 1:
            # weak: every rank holds a full 4096-scene batch of the same global layout
            wms, _ = time_batch(rank * args.batched_scenes, args.batched_scenes, world * args.batched_scenes)
        else:
            wms = rms
        out["batched_scenes"]["weak"] = {"scenes_per_gpu": args.batched_scenes, "scenes_total": world * args.batched_scenes, "ms_per_step": wms / args.steps,
                                         "value": world * args.batched_scenes * args.steps / (wms * 1e-3), "unit": "scene-steps/s", "scaling": "weak"}

    # ---- the other BASELINE.json configurations at full size (device-resident ms/step; parity for them is tests/test_gpu_fullsize.py) ----
    if rank == 0 and world == 1 and args.other_configs:
        other = {}
        for key, mk, settle in (("C1_pyramid_1k_boxes_8_substeps", lambda: S.pyramid(1000), 60), ("C2_mixed_bin_100k", lambda: S.mixed_bin(100_000), 150),
                                ("C3_convex_pile_250k", lambda: S.convex_pile(250_000), 100)):
            try:
                od = mk()
                octx = Context(od, device=local_rank, max_pairs=32 * od.n + 4096, max_manifolds=12 * od.n + 4096)
                for _ in range(settle):
                    octx.step()
                octx.sync()
                ostream = torch.cuda.ExternalStream(octx.stream_ptr(), device=torch.device("cuda", local_rank))
                o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ks = 50
                o0.record(ostream)
                for _ in range(ks):
                    octx.step()
                o1.record(ostream)
                torch.cuda.synchronize()
                oc = octx.counts()
                other[key] = {"ms_per_step": o0.elapsed_time(o1) / ks, "body_steps_per_s": octx.n_dyn * ks / (o0.elapsed_time(o1) * 1e-3), "bodies": int(octx.n_dyn),
                              "pairs": int(oc.n_pairs), "manifolds": int(oc.n_manifolds), "colors": int(oc.n_colors), "substeps": od.substeps,
                              "settle_steps": settle, "islands": octx.island_stats()}
                octx.close()
            except Exception as e:
                other[key] = {"unavailable": str(e)[:200]}
        out["other_configs"] = other

    # ---- the same workload through the host C++ layer: physecs::Scene over an entt::registry (rank 0, N = 1) ------------------
    if rank == 0 and world == 1 and args.scene_bodies > 0:
        try:
            from physecs_b200 import scene_api
            sd = desc if args.scene_bodies >= n else S.terrain(args.scene_bodies, cells=max(16, int(args.scene_bodies ** 0.5 * 1.024)), drop=0.3)
            threads = max((os.cpu_count() or 1) - 1, 0)
            hs = scene_api.HostScene(sd, num_threads=threads, device=local_rank)
            hs.set_arena_capacity(8 * sd.n + 4096, 6 * sd.n + 4096)
            for _ in range(args.settle + args.warmup):
                hs.simulate()
            tt = time.perf_counter()
            ks = min(args.steps, 30)
            for _ in range(ks):
                hs.simulate()
            wall = (time.perf_counter() - tt) / ks
            st = hs.stats()
            out["e2e_scene"] = {"api": "physecs::Scene::simulate over entt::registry (host gather -> pb_set_state -> pb_step -> pb_get_state -> scatter)",
                                "value": sd.n_dynamic / wall, "unit": UNIT, "ms_per_step": wall * 1e3, "bodies": sd.n_dynamic, "host_threads": threads + 1,
                                "gather_ms": st["gather_ms"], "scatter_ms": st["scatter_ms"], "device_ms": st["device_ms"]}
            hs.close()
        except Exception as e:
            out["e2e_scene"] = {"unavailable": str(e)[:200]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            v, ms, nm = reference_sample(args.ref_bodies, args.ref_cells, 100, 3, 15, max(cores - 1, 0))
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "ms_per_step": ms,
                                   "sample": f"reference+hash-fix build, terrain scene with {args.ref_bodies} bodies, settled 100 steps, 15 timed steps, Scene(registry, {max(cores - 1, 0)})"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    else:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "rank 0 at N=1 only"}

    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

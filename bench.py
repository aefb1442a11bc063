#!/usr/bin/env python3
"""bench.py -- headline benchmarks of the Physecs per-step pipeline on B200.

BASELINE.json's metric has two halves, and the line printed follows the half that applies to the launch:

  N = 1   "body-steps/sec & ms/step at 1M bodies" -- workload = BASELINE.json configs[3]: 1,000,000 spheres and capsules over a
          static 2,097,152-triangle terrain mesh, 60 Hz, 4 TGS substeps x 2 iterations (+ relaxation), settled for --settle steps
          so the contact graph is populated.  A "step" is one physecs::Scene::simulate(1/60): broadphase -> narrowphase ->
          contact build -> substep solve -> bounds refresh.  value = body-steps/s with the scene resident in HBM.
  N > 1   "batched-scene steps/s at 1/2/4/8 GPU" -- workload = configs[4]: 4096 independent ragdoll scenes (11 bodies + 10 joints +
          ground each) sharded by scene over the N ranks (contiguous blocks, no collective): value = scene-steps/s of the whole
          batch, "scaling": "strong".  Rank 0 also times the whole batch alone on its GPU in the same run (`one_gpu_same_run`), the
          basis of the scaling ratio; the 1 M-body replicas and the weak batch (4096 scenes per GPU) are extra keys.

Keys of the line (rank 0, one JSON object): value / ms_per_step / p95; e2e = the same metric through the C ABI with HOST buffers
(pb_set_state H2D + pb_step + pb_get_state D2H every step); roofline = the substep loop's kernels (k_integrate_v, k_contact_prep,
k_substep_solve) against MEASURED_PEAKS.json hbm_gbs, with `stages` = every stage of the step on the same terms; cpu_baseline =
the reference's own CPU code (oracle/_ref) on bounded samples, several rows; clocks; gpu_launches.

`--impl reference` times the reference CPU implementation itself on the host cores (same metric, bounded sample).
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
METRIC_BATCH = "batched-scene steps/s (4096 independent ragdoll scenes: 11 bodies + 10 joints + ground each, 4 substeps, sharded by scene)"
UNIT_BATCH = "scene-steps/s"


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


# ---- the reference's own CPU implementation (oracle/_ref), bounded samples ---------------------------------------------------------
def ref_time(desc, threads, hashfix, settle, warmup, steps, presort=True):
    """ms/step list of `steps` Scene::simulate calls of the reference on `desc` after settle + warmup steps."""
    from oracle.ref import RefScene
    ref = RefScene(desc, threads, hashfix=hashfix)
    if presort:
        ref.presort()      # the order the reference's own first-step insertion sort would reach, without its O(n^2) first step
    for _ in range(settle + warmup):
        ref.simulate()
    ms = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref.simulate()
        ms.append((time.perf_counter() - t0) * 1e3)
    nm = len(ref.manifold_keys())
    ref.close()
    return ms, nm


def reference_sample(n_sample, cells, settle, warmup, steps, threads):
    """The C4 workload at reduced size on the reference (hash-fixed build: as shipped it is O(n^2) in contacts)."""
    from physecs_b200 import scenes as S
    d = S.terrain(n_sample, cells=cells, drop=0.3)
    ms, nm = ref_time(d, threads, True, settle, warmup, steps)
    return n_sample * len(ms) / (sum(ms) * 1e-3), float(np.mean(ms)), nm


def _ragdoll_proc(args):
    first, count, total, settle, steps = args
    from physecs_b200 import scenes as S
    d = S.ragdolls(count, seed=0xC5, first_scene=first, total_scenes=total)
    ms, _ = ref_time(d, 0, True, settle, 0, steps, presort=False)
    return sum(ms) * 1e-3


def reference_ragdolls(n_scenes, settle, steps, procs):
    """configs[4] on the reference: independent scenes are embarrassingly parallel on the CPU too, so `procs` processes simulate
    n_scenes / procs scenes each (one reference Scene per process, numThreads = 0); scene-steps/s of the whole set."""
    import multiprocessing as mp
    per = max(1, n_scenes // procs)
    jobs = [(k * per, per, per * procs, settle, steps) for k in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        secs = pool.map(_ragdoll_proc, jobs)
        wall = time.perf_counter() - t0
    return per * procs * steps / max(secs), max(secs) / steps * 1e3, per * procs, wall


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    try:
        if args.gpus > 1:
            procs = min(cores, 32)
            scenes = procs * args.ref_scenes_per_proc
            value, ms, scenes, _ = reference_ragdolls(scenes, args.ref_settle_batch, max(args.steps // 4, 5), procs)
            sample = f"{scenes} of the 4096 ragdoll scenes, {procs} processes x {scenes // procs} scenes (one reference Scene each, numThreads 0), hash-fixed build, settled {args.ref_settle_batch} steps"
            out = {"impl": "reference", "metric": METRIC_BATCH, "value": value, "unit": UNIT_BATCH, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                   "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                   "config": {"workload": "C5: independent ragdoll scenes (revolute / spherical / universal joints + contacts), 60 Hz, 4 substeps x 2 iterations", "sample_scenes": scenes},
                   "cpu_baseline": {"value": value, "unit": UNIT_BATCH, "cores": procs, "kind": "reference", "sample": sample},
                   "e2e": {"value": value, "unit": UNIT_BATCH, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
            print(json.dumps(out))
            return
        threads = max(cores - 1, 0)
        n_sample, cells = args.ref_bodies, args.ref_cells
        value, ms, nm = reference_sample(n_sample, cells, args.ref_settle, args.warmup, max(min(args.steps, 30), 5), threads)
    except Exception as e:  # oracle not built
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref not usable: {e}"}))
        return
    sample = f"terrain scene, {n_sample} bodies on a {cells}x{cells}-cell mesh, settled {args.ref_settle} steps, reference+hash-fix build, Scene(registry, {threads})"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 terrain: spheres+capsules over static triangle mesh, 60 Hz, 4 substeps x 2 iterations", "sample_bodies": n_sample,
                   "manifolds": nm, "note": "same workload at reduced size: the ratio to the device arm is per body-step (the reference needs ~10 s per step at 1 M bodies)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads + 1, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def cpu_baseline_rows(args):
    """SURVEY.md §8d's CPU rows, each a bounded sample: (1) C4 hash-fixed, all cores (the headline row); (2) C4 AS SHIPPED at 4 000 bodies;
    (3) C2 hash-fixed at 100 k bodies (full size); (4) C1 as shipped at FULL size (the one config reference and device run identically)."""
    from physecs_b200 import scenes as S
    cores = os.cpu_count() or 1
    threads = max(cores - 1, 0)
    rows = {}

    def row(key, desc, hashfix, settle, steps, label, presort=True):
        try:
            ms, nm = ref_time(desc, threads, hashfix, settle, 2, steps, presort)
            rows[key] = {"value": desc.n_dynamic / (np.mean(ms) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ms)), "p95_ms": float(np.percentile(ms, 95)),
                         "bodies": desc.n_dynamic, "manifolds": nm, "cores": cores, "kind": "reference", "sample": label + f", settled {settle} steps, {steps} timed steps, Scene(registry, {threads})"}
        except Exception as e:
            rows[key] = {"value": None, "sample": f"unavailable: {e}"}

    row("C4_hashfix_%d" % args.ref_bodies, S.terrain(args.ref_bodies, cells=args.ref_cells, drop=0.3), True, 100, 15, "reference + hash fix, C4 terrain at reduced size")
    if args.cpu_rows:
        row("C4_as_shipped_4000", S.terrain(4000, cells=72, drop=0.3), False, 30, 6, "reference AS SHIPPED (degenerate contact-cache hash), C4 terrain at 4 000 bodies")
        row("C2_hashfix_100k", S.mixed_bin(100_000), True, 20, 4, "reference + hash fix, C2 mixed bin at FULL size (100 k bodies; short settle: the pile is still forming)")
        row("C1_as_shipped_full", S.pyramid(1000), False, 60, 40, "reference AS SHIPPED, C1 1 000-box pyramid at full size (8 substeps x 2 iterations)", presort=False)
    return rows


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
    ap.add_argument("--ref-scenes-per-proc", type=int, default=8)
    ap.add_argument("--ref-settle-batch", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=1, help="also the as-shipped / full-size CPU rows of SURVEY.md 8d (0 = only the headline row)")
    ap.add_argument("--batched-scenes", type=int, default=4096, help="ragdoll scenes of the sharded batch (0 = skip)")
    ap.add_argument("--scene-bodies", type=int, default=1_000_000, help="bodies for the physecs::Scene end-to-end section (0 = skip)")
    ap.add_argument("--other-configs", type=int, default=1, help="also time BASELINE.json's C1 / C2 / C3 scenes at full size (0 = skip)")
    ap.add_argument("--replicas", type=int, default=1, help="N > 1: also time one 1 M-body replica per GPU (0 = skip)")
    ap.add_argument("--ncu", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (never a bench value)")
    ap.add_argument("--ncu-config", default="C4", help="with --ncu: which scene the bracketed region steps (C4, C1, C2, C3, ragdolls512, ragdolls4096)")
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
    from physecs_b200 import batch as B
    from physecs_b200.capi import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the device path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(ctx, steps):
        """`steps` x pb_step on a resident scene: total device ms between two events on the context's stream (max over ranks) and
        the per-step durations (an event after every step, read after the loop: no sync inside)."""
        stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier(); torch.cuda.synchronize()
        ev[0].record(stream)
        for k in range(steps):
            ctx.step()
            ev[k + 1].record(stream)
        barrier(); torch.cuda.synchronize()
        per = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)])
        return max_over_ranks(ev[0].elapsed_time(ev[steps])), per

    peak, peak_src = peaks()

    def pcie_probe(mb=64):
        """pinned <-> device copy rates of this box (GB/s): what the end-to-end figures below are made of"""
        n = mb * 1024 * 1024 // 4
        h = torch.empty(n, dtype=torch.float32).pin_memory()
        dbuf = torch.empty(n, dtype=torch.float32, device=dev)
        out = {}
        for name, src, dst in (("h2d_gbs", h, dbuf), ("d2h_gbs", dbuf, h)):
            best = 0.0
            for _ in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); dst.copy_(src, non_blocking=True); b.record(); torch.cuda.synchronize()
                best = max(best, n * 4 / (a.elapsed_time(b) * 1e-3) / 1e9)
            out[name] = best
        return out

    # ---- scenes of the ncu captures other than the headline one ------------------------------------------------------------------
    if args.ncu and args.ncu_config != "C4":
        mk = {"C1": lambda: S.pyramid(1000), "C2": lambda: S.mixed_bin(100_000), "C3": lambda: S.convex_pile(250_000),
              "ragdolls512": lambda: S.ragdolls(512, total_scenes=4096), "ragdolls4096": lambda: S.ragdolls(4096)}[args.ncu_config]
        d = mk()
        c = Context(d, device=local_rank, max_pairs=32 * d.n + 4096, max_manifolds=12 * d.n + 4096)
        for _ in range(args.settle):
            c.step()
        c.sync()
        c.lib.pb_profiler_range(1)
        for _ in range(args.steps):
            c.step()
        c.sync()
        c.lib.pb_profiler_range(0)
        print(json.dumps({"ncu_capture": args.ncu_config, "steps": args.steps}))
        return

    # =================================================================================================================================
    # batched independent scenes (BASELINE.json configs[4]) through the pb_batch_* C ABI
    # =================================================================================================================================
    def time_batch(first, count, total, e2e=False):
        rd = S.ragdolls(count, seed=0xC5, first_scene=first, total_scenes=total)
        bt = B.Batch([rd], [local_rank])
        rctx = bt.shards[0]
        bt.step(120); bt.sync()
        stream = torch.cuda.ExternalStream(rctx.stream_ptr(), device=dev)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = rctx.launches()
        barrier(); torch.cuda.synchronize()
        r0.record(stream)
        bt.step(args.steps)
        bt.sync()
        r1.record(stream)
        barrier(); torch.cuda.synchronize()
        ms = max_over_ranks(r0.elapsed_time(r1))
        launches = rctx.launches() - l0
        res = {"ms": ms, "bodies": int(rctx.n_dyn), "manifolds": int(rctx.counts().n_manifolds), "joints": len(rd.joints), "islands": rctx.island_stats(), "launches": launches}
        if e2e:
            # host buffers: H2D of every body's state, the step, D2H of the result, every step, through pb_batch_set_state / _step / _get_state
            n = rctx.n_dyn
            bufs = [torch.empty((n, w), dtype=torch.float32).pin_memory() for w in (3, 4, 3, 3)]
            arrs = [[b.numpy()] for b in bufs]
            bt.get_state(*arrs)
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                bt.set_state(*arrs)
                bt.step(1)
                bt.get_state(*arrs)
            torch.cuda.synchronize(); barrier()
            res["e2e_s"] = max_over_ranks(time.perf_counter() - t0)
            res["e2e_bytes"] = 13 * 4 * n
            res["checksum"] = float(np.abs(arrs[0][0]).sum())
        bt.close()
        return res

    batch_out = None
    if args.batched_scenes > 0:
        ns = args.batched_scenes
        b0, b1 = B.shard_range(ns, world, rank)
        sampler_b = ClockSampler(local_rank)
        sampler_b.start()
        r = time_batch(b0, b1 - b0, ns, e2e=True)
        clocks_b = sampler_b.stop()
        batch_out = {"metric": METRIC_BATCH, "value": ns * args.steps / (r["ms"] * 1e-3), "unit": UNIT_BATCH, "scenes": ns, "scenes_on_rank0": b1 - b0,
                     "ms_per_step": r["ms"] / args.steps, "scaling": "strong", "sharding": "contiguous blocks of scenes per rank (pb_batch_shard_range), one context + host thread + stream per GPU, no collective",
                     "bodies_rank0": r["bodies"], "manifolds_last_step_rank0": r["manifolds"], "joints_rank0": r["joints"], "islands": r["islands"], "gpu_launches": int(r["launches"]),
                     "e2e": {"value": ns * args.steps / r["e2e_s"], "unit": UNIT_BATCH, "ms_per_step": r["e2e_s"] / args.steps * 1e3,
                             "h2d_bytes_per_step": r["e2e_bytes"], "d2h_bytes_per_step": r["e2e_bytes"], "checksum": r["checksum"]}, "clocks": clocks_b}
        if world > 1:
            # the basis of the strong-scaling ratio, in the same run: rank 0 alone holds the whole batch (the other ranks wait at the barrier)
            solo = time_batch_solo(S, B, Context, torch, dev, local_rank, ns, args.steps) if rank == 0 else None
            barrier()
            if solo is not None:
                batch_out["one_gpu_same_run"] = {"value": ns * args.steps / (solo * 1e-3), "unit": UNIT_BATCH, "ms_per_step": solo / args.steps, "scenes": ns}
            w = time_batch(rank * ns, ns, world * ns)
            batch_out["weak"] = {"scenes_per_gpu": ns, "scenes_total": world * ns, "ms_per_step": w["ms"] / args.steps,
                                 "value": world * ns * args.steps / (w["ms"] * 1e-3), "unit": UNIT_BATCH, "scaling": "weak"}
        else:
            # one GPU's share of an 8-way split of the same batch: the latency floor that bounds strong scaling
            share = time_batch(0, ns // 8, ns)
            batch_out["share_of_8"] = {"scenes": ns // 8, "ms_per_step": share["ms"] / args.steps, "value": (ns // 8) * args.steps / (share["ms"] * 1e-3), "unit": UNIT_BATCH,
                                       "gpu_launches": int(share["launches"]), "note": "what each GPU of an 8-way strong split steps: bounds the N = 8 value at 8 x this"}

    # =================================================================================================================================
    # N > 1: the batch is the headline; the 1 M-body replicas are an extra key
    # =================================================================================================================================
    if world > 1:
        out = {"metric": METRIC_BATCH, "value": batch_out["value"], "unit": UNIT_BATCH, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": batch_out["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "C5: %d independent ragdoll scenes (revolute / spherical / universal joints + contacts), 60 Hz, 4 substeps x 2 iterations, sharded by scene over %d GPUs" % (args.batched_scenes, world),
                          "scenes": args.batched_scenes, "scenes_per_gpu": batch_out["scenes_on_rank0"], "settle_steps": 120,
                          "l2_policy": "every step rewrites its arenas (pairs, manifolds, rows): no reuse of a previous step's lines is possible"},
               "clocks": batch_out["clocks"], "e2e": batch_out["e2e"], "gpu_launches": batch_out["gpu_launches"], "batched_scenes": batch_out,
               "roofline": {"bound": "latency", "note": "a 512..2048-scene step is bound by dependent kernel launches and colour phases, not by bandwidth: "
                                                        "the HBM roofline of the solver kernels is reported by the N = 1 line on the 1 M-body scene"},
               "cpu_baseline": {"value": None, "unit": UNIT_BATCH, "cores": 0, "kind": "reference", "sample": "rank 0 at N=1 only (bench.py --impl reference --gpus N times the batch on the CPU)"}}
        if args.replicas:
            n = args.bodies
            desc = S.terrain(n, cells=args.cells, drop=0.3)
            ctx = Context(desc, device=local_rank, max_pairs=8 * n + 4096, max_manifolds=6 * n + 4096)
            for _ in range(args.settle + args.warmup):
                ctx.step()
            ctx.sync()
            ms_total, per = timed_steps(ctx, min(args.steps, 50))
            out["replicas_1M"] = {"metric": METRIC, "value": world * ctx.n_dyn * len(per) / (ms_total * 1e-3), "unit": UNIT, "ms_per_step": ms_total / len(per), "scaling": "weak",
                                  "note": "one independent 1 M-body scene per GPU (a single scene does not shard: DESIGN.md multi-GPU)"}
            ctx.close()
        if rank == 0:
            print(json.dumps(out))
        dist.destroy_process_group()
        return

    # =================================================================================================================================
    # N = 1: the 1 M-body scene
    # =================================================================================================================================
    n = args.bodies
    desc = S.terrain(n, cells=args.cells, drop=0.3)
    ctx = Context(desc, device=local_rank, max_pairs=8 * n + 4096, max_manifolds=6 * n + 4096)
    n_dyn = ctx.n_dyn
    for _ in range(args.settle):
        ctx.step()
    ctx.sync()
    for _ in range(args.warmup):
        ctx.step()
    ctx.sync()

    # ---- timed region A: device-resident ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launches()
    sampler.start()
    if args.ncu:
        ctx.lib.pb_profiler_range(1)
    ms_total, per_step = timed_steps(ctx, args.steps)       # no read-back inside the timed loop: the counters stay on the device
    if args.ncu:
        ctx.lib.pb_profiler_range(0)
    clocks = sampler.stop()
    launches = ctx.launches() - launches0
    c = ctx.counts()     # the scene is settled: the last step's counts stand for the timed region
    avgM, avgP, avgC, avgPairs = float(c.n_manifolds), float(c.n_points), max(float(c.n_colors), 1.0), float(c.n_pairs)
    ms_per_step = ms_total / args.steps
    value = world * n_dyn * args.steps / (ms_total * 1e-3)

    # ---- stage times: CUDA events around each stage of pb_step (the context records them every step), read back over a few
    # extra steps OUTSIDE the timed region; the in-kernel phase stamps of the persistent solver likewise
    stage = {"broadphase": [], "narrowphase": [], "contact_build": [], "solve": [], "solve_kernels": []}
    for _ in range(min(args.steps, 20)):
        ctx.step()
        t = ctx.timings()
        stage["broadphase"].append(t.broadphase); stage["narrowphase"].append(t.narrowphase); stage["contact_build"].append(t.contact_build)
        stage["solve"].append(t.solve); stage["solve_kernels"].append(t.solve_kernel)
    st = {k: float(np.mean(v)) for k, v in stage.items()}
    prof_steps = min(args.steps, 20)
    ctx.set_profile(True)
    for _ in range(prof_steps):
        ctx.step()
    prof = ctx.profile()
    ctx.set_profile(False)
    ksolve_ms, ksolve_launches = prof.pop("k_substep_solve_last_step", (0.0, 0))     # CUDA events around each launch of the last profiled step
    phase_ms = {k: v[0] / prof_steps for k, v in prof.items()}
    islands = ctx.island_stats()
    bins = ctx.bin_counts()

    # ---- roofline: algorithmic bytes (SURVEY.md 8d per-unit figures x units) / CUDA-event time / measured HBM peak --------------------
    S_, I_ = desc.substeps, desc.iterations
    n_col = len(desc.col_type)
    bytes_bodies = 260.0 * n_dyn * S_                                   # integrate v 140 + integrate x 120 B / body / substep
    bytes_prep = (244.0 * avgM + 156.0 * avgP) * S_                     # contact prep
    bytes_solve = (180.0 * avgM + 140.0 * avgP) * (I_ + 1) * S_         # contact solve passes (iterations + relaxation)
    bytes_loop = bytes_bodies + bytes_prep + bytes_solve
    bytes_broad = (76.0 + 64.0 + 96.0) * n_col + 24.0 * n_dyn + 8.0 * avgPairs
    bytes_narrow = 112.0 * avgPairs + 24.0 * avgP
    bytes_build = (16.0 + 16.0 + 24.0 * avgP / max(avgM, 1.0)) * avgM + (64.0 + 28.0 * avgP / max(avgM, 1.0)) * avgM   # read raw manifold (key, normal, points) + write constraint header / arms
    loop_ms = st["solve_kernels"]
    achieved = bytes_loop / (loop_ms * 1e-3) / 1e9 if loop_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    tj = {}
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("substep_loop_bytes_per_step", tj.get("k_substeps_bytes_per_launch"))
        except Exception:
            traffic = None

    def stage_row(ms, nbytes, kernels, key):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"ms": ms, "share_of_step": ms / ms_per_step, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac": gbs / peak, "kernels": kernels,
                "traffic": tj.get("stages", {}).get(key)}
    stages = {
        "broadphase": stage_row(st["broadphase"], bytes_broad, ["k_scene_bounds", "k_morton", "k_radix_sort_coop", "k_lbvh_build", "k_lbvh_refit", "k_lbvh_pairs"], "broadphase"),
        "narrowphase": stage_row(st["narrowphase"], bytes_narrow, ["k_world_pose", "k_pair_classify", "k_pair_scatter", "k_np_prim<bin>", "k_np_mesh_light<sphere|capsule>", "k_np_mesh_spill"], "narrowphase"),
        "contact_build": stage_row(st["contact_build"], bytes_build, ["k_color", "k_island_*", "k_scatter_by_key", "k_contact_build"], "contact_build"),
        "substep_loop": stage_row(loop_ms, bytes_loop, ["k_integrate_v", "k_contact_prep", "k_substep_solve"], "substep_loop"),
    }
    stages["narrowphase"]["bins"] = bins
    stages["narrowphase"]["per_bin_ncu"] = tj.get("narrowphase_bins")      # time, DRAM bytes, threads per instruction per bin kernel (ncu capture named in traffic.json)
    pass_ms = phase_ms.get("contact_pass", 0.0) + phase_ms.get("local_sweeps", 0.0)
    # the dominant kernel: k_substep_solve, one launch per substep = every contact colour of (iterations + 1) passes + position integration
    k_launch_ms = ksolve_ms / ksolve_launches if ksolve_launches else 0.0
    k_bytes = (bytes_solve + 120.0 * n_dyn * S_) / S_                      # algorithmic bytes of ONE launch (solve passes + integrate-x)
    k_achieved = k_bytes / (k_launch_ms * 1e-3) / 1e9 if k_launch_ms > 0 else 0.0
    k_traffic = (tj.get("kernels", {}).get("k_substep_solve", {}) or {}).get("dram_bytes_per_step")
    k_traffic = k_traffic / S_ if k_traffic else None
    roofline = {"bound": "hbm", "kernel": "k_substep_solve", "achieved": k_achieved, "peak": peak, "unit": "GB/s", "frac": k_achieved / peak,
                "traffic": k_traffic, "peak_source": peak_src, "bytes_per_launch": k_bytes, "launch_ms": k_launch_ms, "launches_per_step": S_,
                "launches_timed": int(ksolve_launches), "share_of_step": k_launch_ms * S_ / ms_per_step,
                "measured_frac": (k_traffic / (k_launch_ms * 1e-3) / 1e9 / peak) if (k_traffic and k_launch_ms > 0) else None,
                "substep_loop": {"kernels": "k_integrate_v + k_contact_prep + k_substep_solve, %d substeps" % S_, "achieved": achieved, "frac": achieved / peak,
                                 "bytes_per_step": bytes_loop, "ms_per_step": loop_ms, "traffic": traffic, "share_of_step": loop_ms / ms_per_step},
                "phases_from_in_kernel_stamps": {"contact_solve_passes_ms": pass_ms, "contact_solve_passes_GB/s": bytes_solve / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else 0.0,
                                                  "all": phase_ms},
                "stages": stages, "islands": islands,
                "note": "achieved = SURVEY.md 8d algorithmic bytes of one k_substep_solve launch (contact solve passes x (iterations + 1) + integrate-x) / its average launch "
                        "duration by CUDA events on the context's stream (the launches of a step a few steps after the timed region); traffic = DRAM bytes ncu measured for "
                        "the same launch (profiles/traffic.json names the capture), measured_frac = traffic / duration / peak; SURVEY's byte model counts stale-velocity "
                        "reads the kernel does not make, so achieved overstates what moves; substep_loop / stages = the same for the whole loop and for every stage of the step"}

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
    e2e_ms = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        rc = lib.pb_set_state(ctx.ctx, n_dyn, fp(pos), fp(quat), fp(vel), fp(ang))
        rc |= lib.pb_step(ctx.ctx, C.c_float(desc.dt), desc.substeps, desc.iterations, C.c_float(desc.gravity))
        rc |= lib.pb_get_state(ctx.ctx, fp(pos), fp(quat), fp(vel), fp(ang))
        if rc:
            raise RuntimeError(lib.pb_last_error(ctx.ctx).decode())
        e2e_ms.append((time.perf_counter() - ts) * 1e3)
    torch.cuda.synchronize(); barrier()
    serial_s = max_over_ranks(time.perf_counter() - t0)
    serial_ms = list(e2e_ms)
    # The same round trip -- every step's inputs ARE the previous step's results, all of them over the bus both ways -- with the
    # transfers placed beside the device's work: the result is read back on the read stream, poses first (pb_get_state_begin), the next
    # step's broadphase is enqueued at once (pb_step_begin: it runs on the bounds the step left behind, not on the new poses), the
    # poses go back up as soon as they have arrived and the narrowphase behind them (pb_step_narrowphase), the velocities follow
    # while it runs, pb_step continues with the contact build.  (Measured with 1 / 2 / 3 / 8 / 16 chunks: 4.87 / 5.1 / 5.0 / 5.4 / 6.1
    # ms -- the per-call latencies of a chunk outweigh what finer overlap of the two directions gains.)
    CH = 1
    first_c = C.c_int(); count_c = C.c_int()
    lib.pb_get_state(ctx.ctx, fp(pos), fp(quat), fp(vel), fp(ang))
    lib.pb_set_readback_order(ctx.ctx, 1)          # poses of every chunk before the velocities: the narrowphase waits for poses only
    NULLF = C.POINTER(C.c_float)()
    barrier(); torch.cuda.synchronize()
    e2e_ms = []
    t0 = time.perf_counter()
    rc = lib.pb_step_begin(ctx.ctx)
    for k in range(e2e_steps):
        ts = time.perf_counter()
        for half in (0, 1):            # poses up first, velocities behind them
            for c in range(CH if k > 0 else 1):
                if k > 0:
                    rc |= (lib.pb_get_state_wait if half else lib.pb_get_state_wait_poses)(ctx.ctx, c, C.byref(first_c), C.byref(count_c))
                    f0, cnt = first_c.value, count_c.value
                else:
                    f0, cnt = 0, n_dyn
                if cnt:
                    off = lambda a, w: C.cast(C.c_void_p(a.ctypes.data + 4 * w * f0), C.POINTER(C.c_float))
                    if half == 0:
                        rc |= lib.pb_set_state_rows(ctx.ctx, f0, cnt, off(pos, 3), off(quat, 4), NULLF, NULLF)
                    else:
                        rc |= lib.pb_set_state_rows(ctx.ctx, f0, cnt, NULLF, NULLF, off(vel, 3), off(ang, 3))
            if half == 0:
                rc |= lib.pb_step_narrowphase(ctx.ctx)      # needs the poses only: runs while the velocities come back up
        rc |= lib.pb_step(ctx.ctx, C.c_float(desc.dt), desc.substeps, desc.iterations, C.c_float(desc.gravity))
        rc |= lib.pb_get_state_begin(ctx.ctx, fp(pos), fp(quat), fp(vel), fp(ang), CH)
        if k + 1 < e2e_steps:
            rc |= lib.pb_step_begin(ctx.ctx)
        if rc:
            raise RuntimeError(lib.pb_last_error(ctx.ctx).decode())
        e2e_ms.append((time.perf_counter() - ts) * 1e3)
    for c in range(CH):
        rc |= lib.pb_get_state_wait(ctx.ctx, c, C.byref(first_c), C.byref(count_c))
    if rc:
        raise RuntimeError(lib.pb_last_error(ctx.ctx).decode())
    torch.cuda.synchronize(); barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    lib.pb_set_readback_order(ctx.ctx, 0)
    e2e_value = world * n_dyn * e2e_steps / e2e_s
    bytes_io = 13 * 4 * n_dyn
    checksum = float(np.abs(pos).sum())

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "p95_ms_per_step": float(np.percentile(per_step, 95)), "max_ms_per_step": float(per_step.max()),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 terrain: %d spheres+capsules over a static %d-triangle mesh, 60 Hz, 4 substeps x 2 iterations + relaxation" % (n_dyn, 2 * args.cells * args.cells),
                   "bodies_per_gpu": n_dyn, "settle_steps": args.settle, "replicas": world, "l2_policy": "working set >> L2 (per-step traffic ~GBs; no flush needed)",
                   "avg_pairs": avgPairs, "avg_manifolds": avgM, "avg_points": avgP, "avg_colors": avgC},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "p95_ms_per_step": float(np.percentile(e2e_ms[1:] or e2e_ms, 95)), "checksum": checksum,
                "calls": "per step: pb_get_state_wait_poses / _wait + pb_set_state_rows for each of %d chunks (the previous step's result goes back up as it arrives, poses first), pb_step_narrowphase between poses and velocities, pb_step, "
                         "pb_get_state_begin, pb_step_begin (the next broadphase beside the read-back); pinned host buffers, every body both ways every step" % CH,
                "serial": {"ms_per_step": serial_s / e2e_steps * 1e3, "p95_ms_per_step": float(np.percentile(serial_ms, 95)), "value": world * n_dyn * e2e_steps / serial_s,
                           "calls": "pb_set_state, pb_step, pb_get_state one after the other (round 1's loop)"}},
        "gpu_launches": int(launches),
        "pcie": pcie_probe(),
        "roofline": roofline,
        "stage_ms_per_step": st,
    }
    ctx.close()
    if batch_out is not None:
        out["batched_scenes"] = batch_out

    # ---- the other BASELINE.json configurations at full size (device-resident ms/step; parity for them is tests/test_gpu_fullsize.py) ----
    if args.other_configs:
        other = {}
        for key, mk, settle in (("C1_pyramid_1k_boxes_8_substeps", lambda: S.pyramid(1000), 60), ("C2_mixed_bin_100k", lambda: S.mixed_bin(100_000), 150),
                                ("C3_convex_pile_250k", lambda: S.convex_pile(250_000), 100)):
            try:
                od = mk()
                octx = Context(od, device=local_rank, max_pairs=32 * od.n + 4096, max_manifolds=12 * od.n + 4096)
                for _ in range(settle):
                    octx.step()
                octx.sync()
                tot, per = timed_steps(octx, 100)
                oc = octx.counts()
                other[key] = {"ms_per_step": tot / len(per), "p95_ms_per_step": float(np.percentile(per, 95)), "body_steps_per_s": octx.n_dyn * len(per) / (tot * 1e-3), "bodies": int(octx.n_dyn),
                              "pairs": int(oc.n_pairs), "manifolds": int(oc.n_manifolds), "colors": int(oc.n_colors), "substeps": od.substeps,
                              "settle_steps": settle, "islands": octx.island_stats(), "cause": int(oc.cause), "spilled_pairs_last_step": int(oc.n_spilled)}
                octx.close()
            except Exception as e:
                other[key] = {"unavailable": str(e)[:300]}
        out["other_configs"] = other

    # ---- the same workload through the host C++ layer: physecs::Scene over an entt::registry ------------------
    if args.scene_bodies > 0:
        try:
            from physecs_b200 import scene_api
            sd = desc if args.scene_bodies >= n else S.terrain(args.scene_bodies, cells=max(16, int(args.scene_bodies ** 0.5 * 1.024)), drop=0.3)
            threads = max((os.cpu_count() or 1) - 1, 0)
            hs = scene_api.HostScene(sd, num_threads=threads, device=local_rank)
            hs.set_arena_capacity(8 * sd.n + 4096, 6 * sd.n + 4096)
            for _ in range(args.settle + args.warmup):
                hs.simulate()
            ks = min(args.steps, 30)
            walls = []
            for _ in range(ks):
                tt = time.perf_counter()
                hs.simulate()
                walls.append(time.perf_counter() - tt)
            wall = float(np.mean(walls))
            stt = hs.stats()
            out["e2e_scene"] = {"api": "physecs::Scene::simulate over entt::registry (pb_step_begin, gather + pb_set_state_rows in chunks behind the broadphase, pb_step, "
                                       "pb_get_state_begin / _wait + scatter chunk by chunk)",
                                "value": sd.n_dynamic / wall, "unit": UNIT, "ms_per_step": wall * 1e3, "p95_ms_per_step": float(np.percentile(walls, 95) * 1e3), "bodies": sd.n_dynamic,
                                "host_threads": threads + 1, "gather_ms": stt["gather_ms"], "scatter_ms": stt["scatter_ms"], "device_ms": stt["device_ms"]}
            # the same with Scene::setSyncMode(SYNC_DEVICE_AUTHORITATIVE): the registry is written every step but only read for bodies the
            # application announced (registry.patch / notifyBodyChanged) -- the mode for applications that let the solver own the state
            hs.set_sync_mode(True)
            for _ in range(5):
                hs.simulate()
            walls = []
            for _ in range(ks):
                tt = time.perf_counter()
                hs.simulate()
                walls.append(time.perf_counter() - tt)
            out["e2e_scene"]["device_authoritative"] = {"ms_per_step": float(np.mean(walls)) * 1e3, "p95_ms_per_step": float(np.percentile(walls, 95) * 1e3),
                                                        "value": sd.n_dynamic / float(np.mean(walls)), "unit": UNIT}
            # a structural edit of the live registry: ONE body spawned, then destroyed again -- the cost of the simulate() call that
            # follows (the device scene description is brought up to date inside it: `prepare_ms`, Scene::getLastStepStats)
            try:
                hs.set_sync_mode(False)
                hs.simulate()
                one = S.dynamic_only(S.mixed_bin(1, spacing=0.8, seed=0x99), lift=(0.0, 60.0, 0.0))
                edit = {}
                for rep in range(2):
                    tt = time.perf_counter()
                    first_new = hs.add_entities(one)
                    hs.simulate()
                    edit.setdefault("spawn_step_ms", []).append((time.perf_counter() - tt) * 1e3)
                    edit.setdefault("spawn_prepare_ms", []).append(hs.stats()["prepare_ms"])
                    tt = time.perf_counter()
                    hs.destroy_entity(first_new)
                    hs.simulate()
                    edit.setdefault("destroy_step_ms", []).append((time.perf_counter() - tt) * 1e3)
                    edit.setdefault("destroy_prepare_ms", []).append(hs.stats()["prepare_ms"])
                out["e2e_scene"]["structural_edit"] = dict({k: float(min(v)) for k, v in edit.items()},
                                                           note="one body spawned into / destroyed in the live 1 M-entity registry; the step that follows, best of 2")
            except Exception as e:
                out["e2e_scene"]["structural_edit"] = {"unavailable": str(e)[:200]}
            hs.close()
        except Exception as e:
            out["e2e_scene"] = {"unavailable": str(e)[:200]}

    if not args.no_cpu_baseline:
        rows = cpu_baseline_rows(args)
        head = rows.get("C4_hashfix_%d" % args.ref_bodies, {})
        out["cpu_baseline"] = {"value": head.get("value"), "unit": UNIT, "cores": head.get("cores", 0), "kind": "reference", "ms_per_step": head.get("ms_per_step"),
                               "sample": head.get("sample", "unavailable"), "rows": rows}
        c1 = rows.get("C1_as_shipped_full", {})
        if c1.get("value") and "other_configs" in out and "ms_per_step" in out["other_configs"].get("C1_pyramid_1k_boxes_8_substeps", {}):
            out["cpu_baseline"]["C1_same_scene"] = {"reference_ms_per_step": c1["ms_per_step"], "device_ms_per_step": out["other_configs"]["C1_pyramid_1k_boxes_8_substeps"]["ms_per_step"],
                                                    "note": "the one configuration the reference runs as shipped at full size: identical scene on both sides"}
    else:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "skipped (--no-cpu-baseline)"}

    print(json.dumps(out))


def time_batch_solo(S, B, Context, torch, dev, local_rank, ns, steps):
    """rank 0 alone: the whole batch on one GPU, no collective inside (the other ranks wait at the caller's barrier)"""
    rd = S.ragdolls(ns, seed=0xC5)
    bt = B.Batch([rd], [local_rank])
    rctx = bt.shards[0]
    bt.step(120); bt.sync()
    stream = torch.cuda.ExternalStream(rctx.stream_ptr(), device=dev)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    r0.record(stream)
    bt.step(steps); bt.sync()
    r1.record(stream)
    torch.cuda.synchronize()
    ms = r0.elapsed_time(r1)
    bt.close()
    return ms


if __name__ == "__main__":
    main()

"""profiles/traffic.json from an ncu launch list of bench.py --ncu (tools/gpu_profile_r02.sh: gpu__time_duration, dram__bytes_read / _write,
registers, occupancy, threads per instruction for EVERY launch of the captured steps): measured DRAM bytes per step and stage, and the
per-bin rows of the narrowphase.  bench.py reads it: roofline.traffic, roofline.stages[*].traffic, stages.narrowphase.per_bin_ncu.
usage: python tools/traffic_from_launches.py profiles/r02/launches_C4_XXX.csv <steps captured> "<source note>" > profiles/traffic.json"""
import csv, json, sys, collections

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
launch = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    if name.startswith("void "):
        name = name[5:]
    launch.setdefault((r[0], name), {})[r[12]] = float(r[14].replace(",", ""))


def stage_of(k):
    if k.startswith(("k_scene_bounds", "k_morton", "k_radix", "k_lbvh", "k_pairs_bruteforce", "k_tile_bounds")):
        return "broadphase"
    if k.startswith(("k_world_pose", "k_pair_", "k_bin_starts", "k_np_", "k_query")):
        return "narrowphase"
    if k.startswith(("k_integrate_v", "k_contact_prep", "k_joint_fill", "k_substep_solve", "k_step_solve_small")):
        return "substep_loop"
    if k.startswith("k_update_bounds") or k.startswith("k_trimesh_bounds"):
        return "bounds_refresh"
    return "contact_build"


kern = collections.OrderedDict()
for (_, name), d in launch.items():
    a = kern.setdefault(name, dict(launches=0, us=0.0, bytes=0.0, tpi=0.0, regs=0, occ=0.0))
    a["launches"] += 1; a["us"] += d.get("gpu__time_duration.sum", 0) / 1e3
    a["bytes"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    a["tpi"] += d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0); a["regs"] = int(d.get("launch__registers_per_thread", 0))
    a["occ"] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)
stages = collections.OrderedDict()
per_kernel = collections.OrderedDict()
for name, a in kern.items():
    st = stage_of(name)
    s = stages.setdefault(st, dict(us_per_step=0.0, bytes_per_step=0.0))
    s["us_per_step"] += a["us"] / steps; s["bytes_per_step"] += a["bytes"] / steps
    per_kernel[name] = {"stage": st, "launches_per_step": a["launches"] / steps, "us_per_step": round(a["us"] / steps, 2), "dram_bytes_per_step": a["bytes"] / steps,
                        "GB/s": round(a["bytes"] / max(a["us"], 1e-9) / 1e3, 1), "threads_per_instruction": round(a["tpi"] / a["launches"], 2), "registers": a["regs"],
                        "warps_active_pct": round(a["occ"] / a["launches"], 1)}
bins = {k: v for k, v in per_kernel.items() if k.startswith("k_np_")}
out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1],
       "note": "ncu times are cold-cache and serialised (compare shares, not absolutes); bytes = dram__bytes_read.sum + dram__bytes_write.sum",
       "substep_loop_bytes_per_step": stages.get("substep_loop", {}).get("bytes_per_step"),
       "stages": {k: v["bytes_per_step"] for k, v in stages.items()}, "stage_us_per_step_under_ncu": {k: round(v["us_per_step"], 1) for k, v in stages.items()},
       "narrowphase_bins": bins, "kernels": per_kernel}
json.dump(out, sys.stdout, indent=1)
print()

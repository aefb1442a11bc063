"""Summarise an ncu launch list (tools/gpu_profile_r02.sh): per kernel launches, time, DRAM bytes, threads per instruction.
usage: python tools/launch_summary.py gpurun_out/launches_X.csv [steps]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
by = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    if name.startswith("void "):
        name = name[5:]
    d = by.setdefault((r[0], name), {})
    d[r[12]] = float(r[14].replace(",", ""))
agg = collections.OrderedDict()
for (_, name), d in by.items():
    a = agg.setdefault(name, dict(n=0, ns=0.0, rd=0.0, wr=0.0, tpi=0.0, regs=0, occ=0.0))
    a["n"] += 1; a["ns"] += d.get("gpu__time_duration.sum", 0); a["rd"] += d.get("dram__bytes_read.sum", 0); a["wr"] += d.get("dram__bytes_write.sum", 0)
    a["tpi"] += d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0); a["regs"] = int(d.get("launch__registers_per_thread", 0)); a["occ"] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)
tot = sum(a["ns"] for a in agg.values())
print(f"{'kernel':44s} {'n/step':>6s} {'us/step':>9s} {'share':>6s} {'MB/step':>9s} {'GB/s':>7s} {'thr/inst':>8s} {'regs':>4s} {'occ%':>5s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    mb = (a["rd"] + a["wr"]) / steps / 1e6
    print(f"{name[:44]:44s} {a['n'] / steps:6.1f} {a['ns'] / steps / 1e3:9.1f} {a['ns'] / tot * 100:5.1f}% {mb:9.1f} {(a['rd'] + a['wr']) / max(a['ns'], 1):7.0f} {a['tpi'] / a['n']:8.1f} {a['regs']:4d} {a['occ'] / a['n']:5.1f}")
print(f"{'total':44s} {sum(a['n'] for a in agg.values()) / steps:6.1f} {tot / steps / 1e3:9.1f}")

"""profiles/traffic.json from an ncu_summary.py CSV of `bench.py --steps 1 --ncu` under `ncu --set full`: measured DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) per launch of the kernels DESIGN.md quotes.
usage: python tools/traffic_from_summary.py profiles/ncu_full_rXX_summary.csv "<source note>" [substeps=4] > profiles/traffic.json"""
import csv, json, sys, collections

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(s):
    v, u = s.split()
    return float(v.replace(",", "")) * UNIT[u]


rows = list(csv.DictReader(open(sys.argv[1])))
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = r["Kernel Name"]
    tot[name] += val(r["dram__bytes_read.sum"]) + val(r["dram__bytes_write.sum"])
    cnt[name] += 1
per = {k: tot[k] / cnt[k] for k in tot}
names = {"k_substep_solve": "k_substep_solve", "k_contact_prep": "k_contact_prep", "k_integrate_v": "k_integrate_v", "k_lbvh_pairs": "k_lbvh_pairs",
         "void k_np_mesh_light<1>": "k_np_mesh_capsule", "void k_np_mesh_light<0>": "k_np_mesh_sphere", "k_contact_build": "k_contact_build",
         "k_radix_sort_coop": "k_radix_sort_coop", "k_lbvh_refit": "k_lbvh_refit"}
out = {"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}
for k, label in names.items():
    if k in per:
        out[label + "_bytes_per_launch"] = per[k]
# the substep loop of one step (bench.py's roofline unit) = substeps x (integrate-v + contact prep + substep solve)
substeps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
out["k_substeps_bytes_per_launch"] = substeps * (per.get("k_substep_solve", 0) + per.get("k_contact_prep", 0) + per.get("k_integrate_v", 0))
json.dump(out, sys.stdout, indent=1)
print()

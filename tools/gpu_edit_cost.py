"""Cost of a structural edit through physecs::Scene: ONE body spawned into / destroyed in a live registry of N bodies on the terrain
(the simulate() call that follows; `prepare_ms` = bringing the device scene description up to date, Scene::getLastStepStats).
    python tools/gpu_edit_cost.py [bodies] [host library, default = the built one]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2:
    os.environ["PHYSECS_SCENE_LIB"] = os.path.abspath(sys.argv[2])
from physecs_b200 import scenes as S
from physecs_b200 import scene_api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = S.terrain(n, cells=max(16, int(n ** 0.5 * 1.024)), drop=0.3)
hs = scene_api.HostScene(d, num_threads=max((os.cpu_count() or 1) - 1, 0))
hs.set_arena_capacity(8 * d.n + 4096, 6 * d.n + 4096)
t = time.perf_counter(); hs.simulate(); first = (time.perf_counter() - t) * 1e3; first_prepare = hs.stats()["prepare_ms"]
for _ in range(12):
    hs.simulate()
steady = []
for _ in range(5):
    t = time.perf_counter(); hs.simulate(); steady.append((time.perf_counter() - t) * 1e3)
one = S.dynamic_only(S.mixed_bin(1, spacing=0.8, seed=0x99), lift=(0.0, 60.0, 0.0))
rows = {"spawn_step_ms": [], "spawn_prepare_ms": [], "destroy_step_ms": [], "destroy_prepare_ms": []}
for rep in range(3):
    t = time.perf_counter(); e = hs.add_entities(one); hs.simulate()
    rows["spawn_step_ms"].append((time.perf_counter() - t) * 1e3); rows["spawn_prepare_ms"].append(hs.stats()["prepare_ms"])
    t = time.perf_counter(); hs.destroy_entity(e); hs.simulate()
    rows["destroy_step_ms"].append((time.perf_counter() - t) * 1e3); rows["destroy_prepare_ms"].append(hs.stats()["prepare_ms"])
print(json.dumps({"lib": os.path.basename(scene_api.LIB_PATH), "bodies": d.n_dynamic, "first_step_ms": round(first, 1), "first_prepare_ms": round(first_prepare, 1),
                  "steady_step_ms": round(min(steady), 2), **{k: [round(x, 1) for x in v] for k, v in rows.items()}}), flush=True)
hs.close()

#!/bin/bash
# (1) the 1 M-body scene through the group-by-group whole-step kernel (experiment), (2) source-level profile of the whole-step kernel on 512 ragdoll scenes
PB_FUSED_LOCAL_MAX=2000000 python bench.py --steps 30 --warmup 5 --other-configs 0 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench_fusedlocal.json 2> gpurun_out/bench_fusedlocal.err
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_step_solve_small -c 1 -o gpurun_out/step_small_512 -f python bench.py --steps 1 --warmup 3 --settle 120 --ncu --ncu-config ragdolls512 > gpurun_out/ncu_full_512.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_fusedlocal.json'))
print('fused-local 1M:', d['ms_per_step'], d['stage_ms_per_step'], d['gpu_launches'])
PY
tail -c 300 gpurun_out/bench_fusedlocal.err; ls -la gpurun_out/*.ncu-rep

#!/bin/bash
python tools/gpu_batch_time.py 512 4096 > gpurun_out/batch5.log 2>&1
python -m pytest tests/test_gpu_gates.py tests/test_gpu_scene.py "tests/test_gpu_fullsize.py::test_full_size_gates[C5_ragdolls_4096]" tests/test_gpu_deterministic.py -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_pass5.log
for G in 444 1776 3552; do
  PB_ISLAND_GROUPS=$G PB_FUSED_LOCAL_MAX=2000000 python bench.py --steps 30 --warmup 5 --other-configs 0 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench_G$G.json 2> gpurun_out/bench_G$G.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_G$G.json'))
print('G=$G fused-local 1M:', d['ms_per_step'], d['stage_ms_per_step'], d['gpu_launches'])" >> gpurun_out/batch5.log
done
PB_ISLAND_GROUPS=1776 python bench.py --steps 30 --warmup 5 --other-configs 0 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench_G1776_sep.json 2> /dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_G1776_sep.json'))
print('G=1776 separate kernels 1M:', d['ms_per_step'], d['stage_ms_per_step'], d['gpu_launches'])" >> gpurun_out/batch5.log
cat gpurun_out/batch5.log; tail -n 4 gpurun_out/t_pass5.log

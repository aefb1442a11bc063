#!/bin/bash
python tools/gpu_batch_time.py 512 4096 > gpurun_out/batch12.log 2>&1
PB_FUSED_NARROW=0 python tools/gpu_batch_time.py 512 >> gpurun_out/batch12.log 2>&1
PB_FUSED_NARROW=100000 python tools/gpu_batch_time.py 4096 >> gpurun_out/batch12.log 2>&1
python -m pytest tests/test_gpu_gates.py tests/test_gpu_longrun.py tests/test_gpu_deterministic.py tests/test_gpu_scene.py tests/test_gpu_fullsize.py -q -m gpu --timeout 1200 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_pass12.log
python bench.py --steps 50 --warmup 5 --other-configs 1 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err
python -c "
import json
d=json.load(open('gpurun_out/bench12.json'))
print('1M:', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), {k:(round(v['ms_per_step'],3) if 'ms_per_step' in v else v) for k,v in d['other_configs'].items()})" >> gpurun_out/batch12.log
cat gpurun_out/batch12.log; tail -n 6 gpurun_out/t_pass12.log

#!/bin/bash
python tools/gpu_ab.py pyramid 1000 60 PB_ISLAND_LOCAL_MAX=1024 PB_ISLAND_LOCAL_MAX=4096,PB_ISLANDS=1 PB_ISLAND_LOCAL_MAX=16384,PB_ISLANDS=1 > gpurun_out/localmax.log 2>&1
python tools/gpu_ab.py mixed 3000 150 PB_ISLAND_LOCAL_MAX=1024 PB_ISLAND_LOCAL_MAX=16384,PB_ISLANDS=1 >> gpurun_out/localmax.log 2>&1
python tools/gpu_ab.py mixed 20000 150 PB_ISLAND_LOCAL_MAX=1024 PB_ISLAND_LOCAL_MAX=16384,PB_ISLANDS=1 >> gpurun_out/localmax.log 2>&1
cat gpurun_out/localmax.log
python bench.py --steps 50 --warmup 5 --other-configs 0 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err
python -c "
import json
d=json.load(open('gpurun_out/bench14.json'))
r=d['roofline']
print('1M:', round(d['ms_per_step'],4), {k:r[k] for k in ('kernel','achieved','frac','traffic','launch_ms','measured_frac','share_of_step')}, r['substep_loop'])"
tail -c 300 gpurun_out/bench14.err

#!/bin/bash
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --other-configs 0 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name 1M:', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'])"
}
run iso1_walk1 PB_MORTON_ISO=1 PB_WALK=1 PB_MESH_SPLIT=0
run iso1_walk0 PB_MORTON_ISO=1 PB_WALK=0 PB_MESH_SPLIT=0
run iso0_walk0 PB_MORTON_ISO=0 PB_WALK=0 PB_MESH_SPLIT=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
PB_MESH_SPLIT=0 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_C4_r02f.csv python bench.py --steps 2 --warmup 3 --ncu --no-cpu-baseline --batched-scenes 0 --scene-bodies 0 --other-configs 0 > /dev/null 2>&1

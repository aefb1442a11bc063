#!/bin/bash
python -m pytest "tests/test_gpu_gates.py::test_three_gates_mesh_three_stage_form" "tests/test_gpu_gates.py::test_three_gates_mesh_light_modes" tests/test_gpu_spill.py::test_big_shapes_on_fine_mesh "tests/test_gpu_fullsize.py::test_full_size_gates[C4_terrain_1M]" -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_pass6.log
for m in 1 0; do
  PB_MESH_SPLIT=$m python bench.py --steps 30 --warmup 5 --other-configs 0 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench_split$m.json 2> gpurun_out/bench_split$m.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_split$m.json'))
print('PB_MESH_SPLIT=$m 1M:', d['ms_per_step'], d['stage_ms_per_step'], d['gpu_launches'])"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_C4_r02c.csv python bench.py --steps 2 --warmup 3 --ncu --no-cpu-baseline --batched-scenes 0 --scene-bodies 0 --other-configs 0 > /dev/null 2>&1
tail -n 6 gpurun_out/t_pass6.log

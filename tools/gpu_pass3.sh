#!/bin/bash
python tools/gpu_batch_time.py 512 4096 64 > gpurun_out/batch3.log 2>&1
PB_NP_FUSE=0 python tools/gpu_batch_time.py 512 >> gpurun_out/batch3.log 2>&1
python -m pytest tests/test_gpu_gates.py tests/test_gpu_scene.py "tests/test_gpu_fullsize.py::test_full_size_gates[C5_ragdolls_4096]" tests/test_gpu_deterministic.py -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_pass3.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ragdolls512_r02b.csv python bench.py --steps 2 --warmup 3 --settle 120 --ncu --ncu-config ragdolls512 > /dev/null 2>&1
cat gpurun_out/batch3.log; tail -n 4 gpurun_out/t_pass3.log

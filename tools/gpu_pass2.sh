#!/bin/bash
# second GPU pass of round 2: the tests touched since pass 1, the launch lists, a bench line
python -m pytest tests/test_gpu_scene.py tests/test_gpu_queries.py tests/test_gpu_spill.py::test_mtd_query_with_a_big_shape "tests/test_gpu_gates.py::test_whole_step_kernel_group_by_group" -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_pass2.log
bash tools/gpu_profile_r02.sh r02a > gpurun_out/profile.log 2>&1
python bench.py --steps 50 --warmup 5 --other-configs 0 --cpu-rows 0 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -n 5 gpurun_out/t_pass2.log; tail -c 400 gpurun_out/bench1.err

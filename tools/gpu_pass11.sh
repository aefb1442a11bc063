#!/bin/bash
python tools/gpu_ab.py mixed 100000 150 PB_TAIL=1 PB_TAIL=0 > gpurun_out/tail.log 2>&1
python tools/gpu_ab.py convex 250000 100 PB_TAIL=1 PB_TAIL=0 >> gpurun_out/tail.log 2>&1
python tools/gpu_ab.py pyramid 1000 60 PB_TAIL=1 PB_TAIL=0 >> gpurun_out/tail.log 2>&1
python -m pytest tests/test_gpu_gates.py tests/test_gpu_batch.py tests/test_gpu_longrun.py tests/test_gpu_deterministic.py "tests/test_gpu_fullsize.py::test_full_size_gates[C2_mixed_bin_100k]" "tests/test_gpu_fullsize.py::test_full_size_gates[C3_convex_pile_250k]" -q -m gpu --timeout 1200 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_pass11.log
cat gpurun_out/tail.log; tail -n 8 gpurun_out/t_pass11.log

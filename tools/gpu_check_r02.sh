#!/bin/bash
# One GPU-box pass of the whole -m gpu suite (split so that a failure in one part does not hide the others) + a short bench.
# Run from the repo root under gpurun: results land in gpurun_out/.
python -m pytest tests/test_gpu_gates.py tests/test_gpu_golden.py tests/test_gpu_scene.py tests/test_gpu_queries.py tests/test_gpu_longrun.py -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_core.log
python -m pytest tests/test_gpu_spill.py tests/test_gpu_deterministic.py -q -m gpu -s --timeout 900 --timeout-method thread 2>&1 | tail -60 > gpurun_out/t_spill.log
python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s --timeout 1200 --timeout-method thread 2>&1 | tail -30 > gpurun_out/t_full.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench0.json 2> gpurun_out/bench0.err
tail -4 gpurun_out/t_core.log gpurun_out/t_spill.log gpurun_out/t_full.log; tail -c 600 gpurun_out/bench0.err

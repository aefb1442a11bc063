#!/bin/bash
python tools/gpu_batch_time.py 512 4096 64 > gpurun_out/batch15.log 2>&1
python tools/gpu_ab.py pyramid 1000 60 "" >> gpurun_out/batch15.log 2>&1
python tools/gpu_ab.py mixed 100000 150 "" >> gpurun_out/batch15.log 2>&1
python -m pytest tests/test_gpu_gates.py tests/test_gpu_scene.py tests/test_gpu_batch.py -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -5 >> gpurun_out/batch15.log
cat gpurun_out/batch15.log

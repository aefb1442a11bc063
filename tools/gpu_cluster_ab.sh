#!/bin/bash
python tools/gpu_ab.py pyramid 1000 60 PB_CLUSTER=1 PB_CLUSTER=0 > gpurun_out/cluster.log 2>&1
python tools/gpu_ab.py mixed 1500 100 PB_CLUSTER=1,PB_ISLANDS=0 PB_CLUSTER=0,PB_ISLANDS=0 >> gpurun_out/cluster.log 2>&1
python -m pytest tests/test_gpu_gates.py tests/test_gpu_scene.py tests/test_gpu_golden.py -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -5 >> gpurun_out/cluster.log
timeout 600 compute-sanitizer --tool synccheck python -c "
import os; os.environ['PB_ISLANDS']='0'
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
c=Context(S.pyramid(200))
for _ in range(5): c.step()
print('cluster synccheck scene done', c.counts().n_manifolds)" 2>&1 | tail -3 >> gpurun_out/cluster.log
cat gpurun_out/cluster.log

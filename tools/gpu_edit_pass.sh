#!/bin/bash
# scene-level parity tests + the cost of a structural edit at 1 M bodies, for two builds of the host layer
python -m pytest tests/test_gpu_scene.py -q -m gpu -x --timeout 300 --timeout-method thread 2>&1 | tail -6
[ -f physecs_b200/lib/libphysecs_b200_scene_before.so ] && python tools/gpu_edit_cost.py 1000000 physecs_b200/lib/libphysecs_b200_scene_before.so 2>&1 | tail -2
python tools/gpu_edit_cost.py 1000000 2>&1 | tail -2

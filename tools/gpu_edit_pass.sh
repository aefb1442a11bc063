#!/bin/bash
# scene-level parity tests + the cost of a structural edit at 1 M bodies (phases of the re-upload on stderr: PB_TRACE_EDIT)
python -m pytest tests/test_gpu_scene.py -q -m gpu -x --timeout 300 --timeout-method thread 2>&1 | tail -6
PB_TRACE_EDIT=1 python tools/gpu_edit_cost.py ${1:-1000000} 2>&1 | tail -9

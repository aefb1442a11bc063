#!/bin/bash
python bench.py --steps 50 --warmup 5 --other-configs 0 --cpu-rows 0 --batched-scenes 0 --scene-bodies 0 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err
python -c "
import json
d=json.load(open('gpurun_out/bench13.json'))
print('1M:', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), d['pcie'])"
python tools/gpu_e2e_split.py > gpurun_out/e2e_split.log 2>&1; cat gpurun_out/e2e_split.log
python -m pytest "tests/test_gpu_gates.py::test_three_gates_mesh_light_modes" "tests/test_gpu_gates.py::test_three_gates" tests/test_gpu_spill.py::test_big_shapes_on_fine_mesh tests/test_gpu_queries.py "tests/test_gpu_fullsize.py::test_full_size_gates[C4_terrain_1M]" -q -m gpu --timeout 900 --timeout-method thread 2>&1 | tail -8

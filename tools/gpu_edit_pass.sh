#!/bin/bash
# check of the structural-edit work: scene-level + golden parity tests, smoke(), a short bench line with e2e_scene.structural_edit
python -m pytest tests/test_gpu_scene.py tests/test_gpu_golden.py -q -m gpu --timeout 300 --timeout-method thread 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py --no-cpu-baseline --other-configs 0 --batched-scenes 0 --steps 30 --warmup 3 > gpurun_out/bench_edit.json 2> gpurun_out/bench_edit.err
python - <<'P'
import json
d = json.loads([l for l in open('gpurun_out/bench_edit.json') if l.startswith('{')][-1])
es = d['e2e_scene']
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'scene', {k: es[k] for k in es if k not in ('api',)})
P

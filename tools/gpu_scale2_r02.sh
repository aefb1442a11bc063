#!/bin/bash
# our bench line at N GPUs, as the driver launches it (the reference arm is CPU-only and unchanged: tools/gpu_scale_r02.sh runs both).  usage: bash tools/gpu_scale2_r02.sh N
N=${1:-8}
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
E=$(date +%s); echo "wall $((E-S)) s"
tail -n 1 gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
b=d['batched_scenes']
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'solo', b.get('one_gpu_same_run'), 'weak', b.get('weak'), 'replicas', d.get('replicas_1M'))"
tail -c 300 gpurun_out/bench_n$N.err

#!/bin/bash
# bench lines at N GPUs (ours + the reference arm), as the driver launches them.  usage: bash tools/gpu_scale_r02.sh N
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_ref.json 2> gpurun_out/bench_n${N}_ref.err
tail -n 1 gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
b=d['batched_scenes']
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'solo', b.get('one_gpu_same_run'), 'weak', b.get('weak'), 'replicas', d.get('replicas_1M'))"
tail -c 300 gpurun_out/bench_n$N.err; tail -n 1 gpurun_out/bench_n${N}_ref.json | head -c 400

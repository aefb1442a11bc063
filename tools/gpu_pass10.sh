#!/bin/bash
# two GPUs: the whole -m gpu suite (on GPU 0; the batch tests use both), the N = 2 bench lines (ours + reference arm), the sanitizer pass
python -m pytest tests -q -m gpu --timeout 1500 --timeout-method thread 2>&1 | tail -25 > gpurun_out/t_pass10.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
bash tools/gpu_sanitize.sh > gpurun_out/sanitize.log 2>&1
tail -n 5 gpurun_out/t_pass10.log; head -c 1500 gpurun_out/bench_n2.json; echo; tail -c 300 gpurun_out/bench_n2.err; head -c 600 gpurun_out/bench_n2_ref.json; cat gpurun_out/sanitize.log

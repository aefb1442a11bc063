#!/bin/bash
# timed full default bench + reference arm (wall clock of each, as the driver sees it)
S=$(date +%s); python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; E=$(date +%s); echo "bench wall $((E-S)) s"
S=$(date +%s); python bench.py --impl reference > gpurun_out/bench_reference_r02.json 2> gpurun_out/bench_reference_r02.err; E=$(date +%s); echo "reference wall $((E-S)) s"
head -c 600 gpurun_out/bench_r02.json; echo; tail -c 300 gpurun_out/bench_r02.err; head -c 400 gpurun_out/bench_reference_r02.json

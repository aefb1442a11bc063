#!/bin/bash
# Final pass on one GPU: whole -m gpu suite, smoke(), bench lines (ours + reference arm), launch lists of
# every configuration, sanitizer.  Outputs in gpurun_out/ (the ones kept are copied to profiles/r02/).
S=$(date +%s); python -m pytest tests -q -m gpu --timeout 1500 --timeout-method thread 2>&1 | tail -15 > gpurun_out/final_tests.log; E=$(date +%s); echo "gpu suite wall $((E-S)) s" >> gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
S=$(date +%s); python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; E=$(date +%s); echo "bench wall $((E-S)) s" >> gpurun_out/final_tests.log
python bench.py --impl reference > gpurun_out/bench_reference_r02.json 2> gpurun_out/bench_reference_r02.err
bash tools/gpu_profile_r02.sh r02 > gpurun_out/profile_final.log 2>&1
bash tools/gpu_sanitize.sh > gpurun_out/sanitize.log 2>&1
tail -n 4 gpurun_out/final_tests.log; tail -n 2 gpurun_out/final_smoke.log | cut -c1-300; head -c 400 gpurun_out/bench_r02.json; echo; tail -c 300 gpurun_out/bench_r02.err; head -c 300 gpurun_out/bench_reference_r02.json; echo; cat gpurun_out/sanitize.log

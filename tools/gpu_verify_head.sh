#!/bin/bash
# last check of a commit: the whole -m gpu suite, smoke(), one default bench line
S=$(date +%s); python -m pytest tests -q -m gpu --timeout 1500 --timeout-method thread 2>&1 | tail -6 > gpurun_out/verify_tests.log; E=$(date +%s); echo "gpu suite wall $((E-S)) s" >> gpurun_out/verify_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/verify_tests.log
S=$(date +%s); python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "bench rc $? wall $(( $(date +%s) - S )) s" >> gpurun_out/verify_tests.log
cat gpurun_out/verify_tests.log; head -c 300 gpurun_out/verify_bench.json

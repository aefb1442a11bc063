#!/bin/bash
# compute-sanitizer over small scenes of every device path (race / memory checking).  Run from the repo root under gpurun;
# logs land in gpurun_out/ (copy into profiles/ to keep).  memcheck: out-of-bounds / misaligned global, shared and local accesses;
# racecheck: shared-memory hazards (the packet walks' stacks and candidate lists, the CTA-local sweeps, the sorts);
# synccheck: divergent barriers.
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log) summaries"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|sanitize scenes done" gpurun_out/sanitize_$tool.log | tail -3
done

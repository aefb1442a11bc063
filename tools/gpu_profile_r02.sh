#!/bin/bash
# ncu launch lists of every BASELINE.json configuration (time, DRAM bytes, registers, occupancy, threads per instruction per launch).
# Run from the repo root under gpurun; CSVs land in gpurun_out/ (copy the ones to keep into profiles/).
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
TAG=${1:-r02a}
for cfg in ragdolls512 ragdolls4096 C1 C2 C3; do
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${cfg}_${TAG}.csv \
      python bench.py --steps 2 --warmup 3 --settle 120 --ncu --ncu-config $cfg > gpurun_out/ncu_${cfg}.log 2>&1
done
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_C4_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --ncu --no-cpu-baseline --batched-scenes 0 --scene-bodies 0 --other-configs 0 > gpurun_out/ncu_C4.log 2>&1
ls -la gpurun_out/*.csv

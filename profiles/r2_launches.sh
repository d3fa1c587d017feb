#!/bin/bash
# round 2: launch list of the default bench command (B200_PROFILING.md recipe): every launch with its device time
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/launches_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_n1_launches.csv $CMD > gpurun_out/launches_ncu.log 2>&1
tail -2 gpurun_out/launches_plain.log | cut -c1-200; wc -l gpurun_out/r2_bench_n1_launches.csv

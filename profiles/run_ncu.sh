#!/bin/bash
# ncu evidence for one bench command (B200_PROFILING.md recipe).  Usage: profiles/run_ncu.sh <tag> [bench args]
# Writes gpurun_out/<tag>_launches.csv (every launch with its device time) and gpurun_out/<tag>_full.ncu-rep
# (--set full of the sub-cycle kernels).  Numbers printed under ncu are never bench values.
set -u
TAG=$1; shift
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_element|k_node|k_sub' -s 40 -c 4 \
    -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log | cut -c1-400
tail -3 gpurun_out/${TAG}_ncu1.log | cut -c1-300
tail -3 gpurun_out/${TAG}_ncu2.log | cut -c1-300
ls -la gpurun_out/

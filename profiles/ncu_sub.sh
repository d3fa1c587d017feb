#!/bin/bash
# full ncu capture (with source) of the sub-cycle kernel on a given workload: profiles/ncu_sub.sh <tag> <workload>
TAG=$1; WL=$2
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_subcycle -s 20 -c 2 -o gpurun_out/${TAG} -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200

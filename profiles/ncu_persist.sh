#!/bin/bash
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
NSX_PERSIST=1 $CMD > gpurun_out/r1f_plain.log 2>&1 &&
NSX_PERSIST=1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_direct_persistent -s 1 -c 1 -o gpurun_out/r1f -f $CMD > gpurun_out/r1f_ncu.log 2>&1
tail -2 gpurun_out/r1f_ncu.log | cut -c1-200

#!/bin/bash
# round 2: ncu full capture of the final k_resident (mailboxes + balanced tiles + slot placement), 10 km mesh
mkdir -p gpurun_out
NSX_PATH=resident timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err_res.log | cut -c1-300
NSX_PATH=resident timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/r2_resident_v4_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -2
ncu -i gpurun_out/r2_resident_v4_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2_resident_v4_raw.csv
ncu -i gpurun_out/r2_resident_v4_full.ncu-rep --page source --csv 2>/dev/null > gpurun_out/r2_resident_v4_source.csv
ls -la gpurun_out/r2_resident_v4*

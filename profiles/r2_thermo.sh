#!/bin/bash
# round 2, GPU call: thermo() parity tests, kernel timing at the 3 km mesh, ncu capture of k_thermo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermo.py -x -q -s 2>&1 | tail -45
timeout 600 python profiles/thermo_bench.py --mesh 3km > gpurun_out/r2_thermo_bench_3km.json 2> gpurun_out/thermo_bench_err.log
tail -3 gpurun_out/thermo_bench_err.log; cat gpurun_out/r2_thermo_bench_3km.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_thermo -s 2 -c 1 -o gpurun_out/r2_thermo_full -f \
    python profiles/thermo_bench.py --mesh 10km --steps 2 --warmup 2 --cpu-elements 1000 2>&1 | tail -2
ncu -i gpurun_out/r2_thermo_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2_thermo_raw.csv
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2_thermo_raw.csv")))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__cycles_active.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"):
        print(k, d.get(k))
PY

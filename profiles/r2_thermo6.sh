#!/bin/bash
# round 2: ncu capture of the final k_thermo (v5: reciprocal reuse), 3 km mesh
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_thermo -s 2 -c 1 -o gpurun_out/r2_thermo_v5_full -f \
    python profiles/thermo_bench.py --mesh 3km --steps 2 --warmup 2 --cpu-elements 1000 2>&1 | tail -1
ncu -i gpurun_out/r2_thermo_v5_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2_thermo_v5_raw.csv
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2_thermo_v5_raw.csv")))
h = rows[0]
d = dict(zip(h, rows[2]))
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "thread_inst_executed_true",
          "smsp__thread_inst_executed_per_inst_executed.ratio"):
    print(k, d.get(k))
for k, v in d.items():
    if "warps_issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v) > 0.1: print("%6.2f %s" % (float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
PY

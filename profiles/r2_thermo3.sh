#!/bin/bash
# round 2: thermo() final shape (128 threads, 4 CTAs/SM), coupled-step parity, 3 km timing with spatially smooth and with
# per-element random fields, ncu capture, and the N=1 bench line with the device-resident step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermo.py -x -q -s 2>&1 | grep -v "^thermo [a-z0-9_]*:" | tail -8
timeout 600 python profiles/thermo_bench.py --mesh 3km --tag smooth > gpurun_out/r2_thermo_bench_3km.json 2> gpurun_out/thermo_bench_err.log
timeout 600 python profiles/thermo_bench.py --mesh 3km --random-regimes --tag random --cpu-elements 1000 > gpurun_out/r2_thermo_bench_3km_random.json 2>> gpurun_out/thermo_bench_err.log
tail -2 gpurun_out/thermo_bench_err.log; cat gpurun_out/r2_thermo_bench_3km.json; cut -c1-700 gpurun_out/r2_thermo_bench_3km_random.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_thermo -s 2 -c 1 -o gpurun_out/r2_thermo_v3_full -f \
    python profiles/thermo_bench.py --mesh 3km --steps 2 --warmup 2 --cpu-elements 1000 2>&1 | tail -2
ncu -i gpurun_out/r2_thermo_v3_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2_thermo_v3_raw.csv
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2_thermo_v3_raw.csv")))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "thread_inst_executed_true",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"):
        print(k, d.get(k))
PY
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_10km_thermo.json 2> gpurun_out/bench_err.log
tail -3 gpurun_out/bench_err.log
python - <<PY
import json
for l in open("gpurun_out/r2_bench_10km_thermo.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], json.dumps(d.get("next_rows"))[:1500])
PY

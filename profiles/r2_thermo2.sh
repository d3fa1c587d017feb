#!/bin/bash
# round 2: thermo() after the instruction-count cuts (integer powers, shared atmosphere terms, reciprocal of uniform
# divisors, register-resident element): parity tests, launch-shape sweep, 3 km timing, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermo.py -x -q 2>&1 | tail -5
for v in t128_m1 t128_m4 t64_m8 t256_m2 t128_m5 t64_m6; do
  NSX_LIBRARY=$PWD/nextsim_b200/_variants/libnsx_$v.so timeout 300 python profiles/thermo_bench.py --mesh 3km --nx 600 --tag $v --cpu-elements 1000 2>>gpurun_out/thermo_sweep_err.log >> gpurun_out/r2_thermo_sweep.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/r2_thermo_sweep.jsonl"):
    d = json.loads(l)
    print(d["tag"], d["regimes"], "%.3f ms" % d["ms_per_call"]["median"], "%.3e el/s" % d["value"], "hbm frac %.3f" % d["roofline"]["frac"])
PY
timeout 600 python profiles/thermo_bench.py --mesh 3km --tag default > gpurun_out/r2_thermo_bench_3km.json 2> gpurun_out/thermo_bench_err.log
timeout 600 python profiles/thermo_bench.py --mesh 3km --nx 600 --random-regimes --tag default_random --cpu-elements 1000 >> gpurun_out/r2_thermo_sweep.jsonl 2>> gpurun_out/thermo_sweep_err.log
tail -2 gpurun_out/thermo_bench_err.log; cat gpurun_out/r2_thermo_bench_3km.json; tail -1 gpurun_out/r2_thermo_sweep.jsonl | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_thermo -s 2 -c 1 -o gpurun_out/r2_thermo_v2_full -f \
    python profiles/thermo_bench.py --mesh 3km --nx 600 --steps 2 --warmup 2 --cpu-elements 1000 2>&1 | tail -2
ncu -i gpurun_out/r2_thermo_v2_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2_thermo_v2_raw.csv
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2_thermo_v2_raw.csv")))
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

#!/bin/bash
# round 2, GPU call 4: resident v2 bench with the in-kernel split + ncu full capture of k_resident
mkdir -p gpurun_out
NSX_PATH=resident timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err_res.log > gpurun_out/bench_10km_resident.json
python - <<PY
import json
for l in open("gpurun_out/bench_10km_resident.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("resident", "%.4g" % d["value"], d["roofline"]["us_per_subcycle"], d["roofline"]["frac"], d["phase_ms"], d["check"], "e2e %.4g" % d["e2e"]["value"])
PY
tail -3 gpurun_out/bench_err_res.log
NSX_PATH=resident ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/r2_resident_v2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -2

#!/bin/bash
# round 2: k_resident checks after a kernel change: parity of the resident path + N=1 bench
mkdir -p gpurun_out

timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "resident" 2>&1 | tail -3
NSX_PATH=resident timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err_res.log > gpurun_out/bench_10km_resident_ow.json
python - <<PY
import json
for l in open("gpurun_out/bench_10km_resident_ow.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("resident", "%.4g" % d["value"], d["roofline"]["us_per_subcycle"], d["phase_ms"], d["check"])
PY

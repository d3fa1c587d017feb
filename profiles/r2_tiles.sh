#!/bin/bash
# round 2: tile path after the 4-shape-plane / displacement-accumulator change: parity subset + 3 km and 1 km benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -k "(tiles or direct) and not 1km and not full_size" 2>&1 | tail -4
for WL in 3km $1; do
  NSX_PATH=tiles timeout 600 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log > gpurun_out/bench_${WL}_tiles.json
  python - <<PY
import json
for l in open("gpurun_out/bench_${WL}_tiles.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$WL tiles", "%.4g" % d["value"], "us/sub %.2f" % d["roofline"]["us_per_subcycle"], "frac %.3f" % d["roofline"]["frac"], d["phase_ms"], d["check"], "e2e %.4g" % d["e2e"]["value"])
PY
  tail -2 gpurun_out/bench_err.log
done

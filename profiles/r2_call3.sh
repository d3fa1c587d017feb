#!/bin/bash
# round 2, GPU call 3: resident solver v2 (flags, pipelined schedule, in-kernel smoother, ranks sharing one GPU)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resident.py -x -q 2>&1 | tail -25
for P in direct resident; do
  NSX_PATH=$P timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err_$P.log > gpurun_out/bench_10km_$P.json
  python - <<PY
import json
for l in open("gpurun_out/bench_10km_$P.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$P", "%.4g" % d["value"], d["roofline"]["us_per_subcycle"], d["roofline"]["frac"], d["phase_ms"], d["check"], "e2e %.4g" % d["e2e"]["value"])
PY
  tail -3 gpurun_out/bench_err_$P.log
done

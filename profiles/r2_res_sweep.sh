#!/bin/bash
# round 2: tiles (CTAs) per SM sweep of the resident solver.  usage: r2_res_sweep.sh "TPB:CTAS ..."
mkdir -p gpurun_out
for cfg in $1; do
  TPB=${cfg%%:*}; CT=${cfg##*:}
  NSX_RES_TPB=$TPB NSX_RES_CTAS=$CT python -c "from nextsim_b200 import build; build.build(force=True)" 2>&1 | tail -2
  echo "== RES_TPB=$TPB RES_CTAS=$CT"
  timeout 300 python -m pytest tests/test_gpu_resident.py -x -q 2>&1 | tail -2
  NSX_PATH=resident timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log > gpurun_out/bench_res_${TPB}_${CT}.json
  python - <<PY
import json
for l in open("gpurun_out/bench_res_${TPB}_${CT}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("resident $cfg", "%.4g" % d["value"], "us/sub %.2f" % d["roofline"]["us_per_subcycle"], d["phase_ms"], d["check"])
PY
  tail -2 gpurun_out/bench_err.log
done

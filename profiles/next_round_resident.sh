#!/bin/bash
# First GPU call of the next round: validate and time the experimental state-resident solver (NSX_PATH=resident).
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash profiles/next_round_resident.sh'
# 1. parity (opt-in tests), 2. bench on the headline mesh with both paths, 3. launch list of the resident run.
set -x
NSX_TEST_RESIDENT=1 python -m pytest tests/test_gpu_resident.py -x -q 2>&1 | tail -15
for P in direct resident; do
  NSX_PATH=$P python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_10km_$P.json
  python - <<PY
import json
for l in open("gpurun_out/bench_10km_$P.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$P", d["value"], d["roofline"]["us_per_subcycle"], d["roofline"]["frac"], d["phase_ms"], d["check"])
PY
done
NSX_PATH=resident ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/resident_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_resident.log 2>&1
grep -c k_resident gpurun_out/resident_launches.csv

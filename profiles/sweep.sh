#!/bin/bash
# quick parameter sweep of the sub-cycle kernel: prints us per sub-cycle for each setting
for spec in "$@"; do
  out=$(env $spec python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
  echo "$spec :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("us/sub=%.2f frac=%.3f value=%.3e e2e=%.3e phases=%s"%(d["roofline"]["us_per_subcycle"], d["roofline"]["frac"], d["value"], d["e2e"]["value"], d["phase_ms"]))' 2>&1 | tail -1)"
done

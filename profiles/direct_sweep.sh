#!/bin/bash
# direct-path kernel shape experiments: "tpb minb"
run() {
  NSX_DIRECT_TPB=$1 NSX_DIRECT_MINB=$2 python -c "from nextsim_b200 import build; build.build(force=True)" 2>&1 | grep -i "spill\|error" | head -3
  out=$(python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
  echo "direct tpb=$1 minb=$2 :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("us/sub=%.2f frac=%.3f value=%.3e"%(d["roofline"]["us_per_subcycle"], d["roofline"]["frac"], d["value"]))' 2>&1 | tail -1)"
}
for spec in "$@"; do run $spec; done
python -c "from nextsim_b200 import build; build.build(force=True)"

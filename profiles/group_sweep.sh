#!/bin/bash
# consumer-group / stage experiments: "stages groups tpb tile"
run() {
  NSX_SUB_STAGES=$1 NSX_SUB_GROUPS=$2 NSX_SUB_TPB=$3 python -c "from nextsim_b200 import build; build.build(force=True)" || return
  for wl in 10km 3km; do
    out=$(NSX_TILE_NODES=$4 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
    echo "stages=$1 groups=$2 tpb=$3 tile=$4 $wl :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("us/sub=%.2f frac=%.3f"%(d["roofline"]["us_per_subcycle"], d["roofline"]["frac"]))' 2>&1 | tail -1)"
  done
}
for spec in "$@"; do run $spec; done
python -c "from nextsim_b200 import build; build.build(force=True)"

#!/bin/bash
# pipeline-depth experiments: rebuild with NSX_SUB_STAGES (and CTA size), time 10km and 3km
run() {  # stages tpb tile
  NSX_SUB_STAGES=$1 NSX_SUB_TPB=$2 python -c "from nextsim_b200 import build; build.build(force=True)" || return
  for wl in 10km 3km; do
    out=$(NSX_TILE_NODES=$3 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
    echo "stages=$1 tpb=$2 tile=$3 $wl :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("us/sub=%.2f frac=%.3f"%(d["roofline"]["us_per_subcycle"], d["roofline"]["frac"]))' 2>&1 | tail -1)"
  done
}
for spec in "$@"; do run $spec; done
python -c "from nextsim_b200 import build; build.build(force=True)"

#!/bin/bash
# kernel-shape experiments: rebuild libnsx.so with a CTA size / min-blocks pair, then time 10km and 3km
run() {  # tpb minb occ tile
  NSX_SUB_TPB=$1 NSX_SUB_MINB=$2 python -c "from nextsim_b200 import build; build.build(force=True)" || return
  for wl in 10km 3km; do
    out=$(NSX_SUB_OCC=$3 NSX_TILE_NODES=$4 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
    echo "tpb=$1 minb=$2 occ=$3 tile=$4 $wl :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("us/sub=%.2f frac=%.3f"%(d["roofline"]["us_per_subcycle"], d["roofline"]["frac"]))' 2>&1 | tail -1)"
  done
}
for spec in "$@"; do run $spec; done
python -c "from nextsim_b200 import build; build.build(force=True)"

#!/bin/bash
# element/node warp role split experiments (tile path): "node_threads tpb stages tile"
run() {
  NSX_SUB_NODE_THREADS=$1 NSX_SUB_TPB=$2 NSX_SUB_STAGES=$3 python -c "from nextsim_b200 import build; build.build(force=True)" || return
  if [ "$5" = "test" ]; then NSX_PATH=tiles timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiles and (stable or partitioned or 3km)" 2>&1 | tail -1; fi
  out=$(NSX_PATH=tiles NSX_TILE_NODES=$4 python bench.py --workload 3km --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
  echo "node_threads=$1 tpb=$2 stages=$3 tile=$4 3km :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("us/sub=%.2f frac=%.3f"%(d["roofline"]["us_per_subcycle"], d["roofline"]["frac"]))' 2>&1 | tail -1)"
}
for spec in "$@"; do run $spec; done
python -c "from nextsim_b200 import build; build.build(force=True)"

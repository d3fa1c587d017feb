#!/bin/bash
# round 2: mailbox exchange on the tile / direct paths.  usage: r2_multi_tiles.sh N
N=$1; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for P in tiles direct; do
  for D in bbm mevp; do
    NSX_PATH=$P timeout 600 $TR tests/run_multigpu_parity.py --nx 128 --dyn $D --steps 2 2>&1 | grep -E "parity|MISMATCH|Error|error|timed out" | head -5
  done
done
for OV in 1 2; do
NSX_OVERLAP=$OV timeout 900 $TR bench.py --gpus $N --workload 3km --scaling strong --steps 5 --warmup 3 --no-north-star --no-parity > gpurun_out/bench_3km_n${N}_ov$OV.json 2> gpurun_out/bench_3km_n$N.err
tail -3 gpurun_out/bench_3km_n$N.err | grep -v OMP
python - <<PY
import json
for l in open("gpurun_out/bench_3km_n${N}_ov$OV.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("3km strong N=$N overlap=$OV", "%.4g" % d["value"], d["config"]["path"], "us/sub %.2f" % d["roofline"]["us_per_subcycle"], d["phase_ms"], d["check"])
PY
done

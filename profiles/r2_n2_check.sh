#!/bin/bash
# round 2: 2-GPU check of the final k_resident (parity preflight + weak scaling + 3 km strong through the tile path)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/bench_n2_err.log
tail -3 gpurun_out/bench_n2_err.log
python - <<PY
import json
for l in open("gpurun_out/r2_bench_n2_final.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=2 value %.4g" % d["value"], "ms/step", d["ms_per_step"], "parity", d.get("parity"), "e2e %.4g" % d["e2e"]["value"])
        for k, v in (d.get("north_star") or {}).items():
            print("  north_star", k, "%.4g" % v["value"], v["config"].get("path"), v.get("check"))
PY

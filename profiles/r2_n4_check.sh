#!/bin/bash
# round 2: 4-GPU bench line of the final kernels (weak scaling at the headline size + 3 km strong)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/bench_n4_err.log
tail -2 gpurun_out/bench_n4_err.log
python - <<PY
import json
for l in open("gpurun_out/r2_bench_n4_final.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=4 value %.4g" % d["value"], "ms/step", d["ms_per_step"], "us/sub", d["roofline"]["us_per_subcycle"], "parity", d.get("parity", {}).get("worst_rel_l2"), "e2e %.4g" % d["e2e"]["value"], d["phase_ms"])
        for k, v in (d.get("north_star") or {}).items():
            print("  north_star", k, "%.4g" % v["value"], v["config"].get("path"), v["phase_ms"])
PY

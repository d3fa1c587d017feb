#!/bin/bash
# round 2: multi-GPU run.  usage: r2_multi.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for P in resident tiles direct; do
  NSX_PATH=$P timeout 600 $TR tests/run_multigpu_parity.py --nx 128 --dyn bbm --steps 2 2>&1 | grep -E "parity|MISMATCH|Error|error" | head -5
done
NSX_PATH=resident timeout 600 $TR tests/run_multigpu_parity.py --nx 160 --dyn mevp --steps 2 2>&1 | grep -E "parity|MISMATCH|Error|error" | head -5
timeout 1500 $TR bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/bench_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l)
        if "value" not in d:
            print(d); continue
        print("N=$N weak", "%.4g" % d["value"], "us/sub %.2f" % d["roofline"]["us_per_subcycle"], d["phase_ms"], d["check"], "parity", d.get("parity"), "clocks", d["clocks"]["samples"], d["clocks"]["reasons"])
        for k, v in (d.get("north_star") or {}).items():
            print("  north_star", k, "%.4g" % v["value"], "ms/step %.3f" % v["ms_per_step"], v["config"]["path"], "us/sub %.2f" % v["roofline"]["us_per_subcycle"], "frac %.3f" % v["roofline"]["frac"], "clock samples", v["clocks"]["samples"], "setup %.0fs" % v["setup_seconds"], "bad", v["check_bad_entries_all_ranks"])
PY

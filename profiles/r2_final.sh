#!/bin/bash
# round 2, final pass: the whole GPU suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -9
timeout 900 python bench.py > gpurun_out/r2_bench_10km_1gpu.json 2> gpurun_out/bench_final_err.log
tail -2 gpurun_out/bench_final_err.log
python - <<PY
import json
for l in open("gpurun_out/r2_bench_10km_1gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "roofline", d["roofline"]["us_per_subcycle"], "e2e %.4g" % d["e2e"]["value"],
              "cpu %.4g x%d" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), "launches", d["gpu_launches"], d["clocks"])
        nr = d["next_rows"]
        print("thermo_us", nr["thermo_us"], "resident_step", nr["resident_step"]["ms_per_step"], nr["resident_step"]["fraction_of_device_rate"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-400

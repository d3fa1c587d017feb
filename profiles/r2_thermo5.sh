#!/bin/bash
# round 2: thermo() with per-element reciprocal reuse; arithmetic variants (FMA contraction on, IEEE divisions everywhere)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermo.py -x -q -s 2>&1 | grep -v "^\.\?thermo [a-z0-9_]*:" | tail -8
rm -f gpurun_out/r2_thermo_sweep2.jsonl
for v in base fmad exactdiv; do
  NSX_LIBRARY=$PWD/nextsim_b200/_variants/libnsx_$v.so timeout 300 python profiles/thermo_bench.py --mesh 3km --tag $v --cpu-elements 1000 2>>gpurun_out/thermo_sweep_err.log >> gpurun_out/r2_thermo_sweep2.jsonl
done
NSX_LIBRARY=$PWD/nextsim_b200/_variants/libnsx_fmad.so timeout 600 python -m pytest tests/test_gpu_thermo.py -x -q -s -k "large_mesh or coupled or golden" 2>&1 | tail -6
python - <<PY
import json
for l in open("gpurun_out/r2_thermo_sweep2.jsonl"):
    d = json.loads(l)
    print(d["tag"], d["regimes"], "%.3f ms" % d["ms_per_call"]["median"], "%.3e el/s" % d["value"], "hbm frac %.3f" % d["roofline"]["frac"])
PY

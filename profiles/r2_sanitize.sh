#!/bin/bash
# round 2: compute-sanitizer over the sub-cycle kernels, the halo kernels and the transfer / diagnostics kernels
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool, tag, args...
  tool=$1; tag=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 10 python profiles/sanitize_case.py "$@" > gpurun_out/san_${tool}_${tag}.log 2>&1
  echo "== $tool $tag (rc $?): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' gpurun_out/san_${tool}_${tag}.log | tail -2 | tr '\n' ' ')  | $(grep -E '^path' gpurun_out/san_${tool}_${tag}.log | tr '\n' ';')"
}
run memcheck tiles1 --path tiles --nranks 1
run memcheck tiles3 --path tiles --nranks 3
run memcheck direct3 --path direct --nranks 3
run memcheck resident1 --path resident --nranks 1
run racecheck tiles1 --path tiles --nranks 1 --substeps 4
run racecheck tiles3 --path tiles --nranks 3 --substeps 3
run racecheck resident1 --path resident --nranks 1 --substeps 4
run synccheck tiles1 --path tiles --nranks 1 --substeps 4
run synccheck resident1 --path resident --nranks 1 --substeps 4
run memcheck resident2 --path resident --nranks 2 --substeps 3

#!/bin/bash
# Line / branch coverage of the dynamics oracle (nextsim_oracle.cpp) under the CPU tests that hold it to the reference's own
# function bodies and to the committed goldens, measured with gcov on an instrumented build (TEST INFRASTRUCTURE ONLY).
#   usage: bash oracle/oracle_coverage.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
cd "$W"
g++ -std=c++17 -O0 -g --coverage -ffp-contract=off -fPIC -shared -pthread -o "$W/liboracle.so" "$ROOT/oracle/nextsim_oracle.cpp"
cp "$ROOT/oracle/_build/liboracle.so" "$W/orig.so"
trap 'cp "$W/orig.so" "$ROOT/oracle/_build/liboracle.so"; touch "$ROOT/oracle/_build/liboracle.so"; rm -rf "$W"' EXIT
cp "$W/liboracle.so" "$ROOT/oracle/_build/liboracle.so"; touch "$ROOT/oracle/_build/liboracle.so"
(cd "$ROOT" && python -m pytest tests/test_ref_fe_cpu.py tests/test_oracle_cpu.py tests/test_oracle_next_rows_cpu.py \
    tests/test_oracle_vs_independent_cpu.py tests/test_partition_cpu.py tests/test_partmesh_cpu.py -q -p no:cacheprovider 2>&1 | tail -1)
gcov -b -o "$W/liboracle.so-nextsim_oracle.gcno" "$ROOT/oracle/nextsim_oracle.cpp" 2>/dev/null | grep -A3 "nextsim_oracle.cpp'" | head -4
echo "lines never executed:"
grep -n "#####" "$W/nextsim_oracle.cpp.gcov" | cut -c1-140

#!/bin/bash
# Line / branch coverage of nsx::thermo::thermo_element() under tests/test_thermo_cpu.py, measured with gcov on an
# instrumented host build (TEST INFRASTRUCTURE ONLY).  The instrumented build is -O0, where gcc does not fold pow(x, 2)
# into x*x as the -O2 builds of both sides do, so the bit-for-bit assertions of the tests are expected to fail here: the
# script only reads which lines ran.   usage: bash oracle/thermo_coverage.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
cd "$W"
g++ -std=c++17 -O0 -g --coverage -ffp-contract=off -fPIC -shared -x c++ -o "$W/libthermo_cov.so" "$ROOT/oracle/thermo_oracle.cpp"
(cd "$ROOT" && THERMO_ORACLE_LIB="$W/libthermo_cov.so" python -m pytest tests/test_thermo_cpu.py -q -p no:cacheprovider > "$W/pytest.log" 2>&1 || true)
tail -1 "$W/pytest.log"
gcov -b -o "$W/libthermo_cov.so-thermo_oracle.gcno" "$ROOT/oracle/thermo_oracle.cpp" 2>/dev/null | grep -A3 "nsx_thermo.cuh" | head -4
echo "lines never executed:"
grep -n "#####" "$W/nsx_thermo.cuh.gcov" | cut -c1-140
rm -rf "$W"

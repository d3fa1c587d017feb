#!/bin/bash
# round 2: compute-sanitizer is closed on this GPU pool ("runs under it have left GPUs needing a reset"); stand-in:
# libnsx.so rebuilt with -DNSX_DEBUG_CHECKS (device-side bounds checks of every index table the sub-cycle kernels trust;
# a violation sets the handle's error word and nsx_download fails) and the parity tests of all three paths run on it.
mkdir -p gpurun_out
/usr/local/cuda/bin/compute-sanitizer --tool memcheck python profiles/sanitize_case.py 2>&1 | head -3
NSX_DEBUG_CHECKS=1 python -c "from nextsim_b200 import build; build.build(force=True)" 2>&1 | tail -2
grep -c "NSX_DEBUG_CHECKS" nextsim_b200/build.log
timeout 1200 python -m pytest tests/test_gpu_resident.py tests/test_gpu_golden.py tests/test_gpu_parity.py -q -m gpu \
   -k "not 3km and not 1km and not full_size" 2>&1 | tail -6
python -c "from nextsim_b200 import build; build.build(force=True)" 2>&1 | tail -2

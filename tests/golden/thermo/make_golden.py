#!/usr/bin/env python
"""Writes tests/golden/thermo/<name>.npz from the REFERENCE's own thermo() bodies (oracle/ref_fe).

Run in the container that holds /root/reference:   python tests/golden/thermo/make_golden.py
Inputs are regenerated from (option set name, nx, seed) by tests/thermo_common.make_inputs, so only the outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))                      # tests/
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))     # repository root

import thermo_common as tc  # noqa: E402

NX, SEED, STEPS = 20, 777, 3

for name in ("defaults", "zero_layer", "alb4_ponds", "nudged_ocean"):
    p, t, dt, gm, S = tc.make_inputs(name, nx=NX, seed=SEED)
    out = tc.run_reference(p, t, dt, gm, S, steps=STEPS)
    path = os.path.join(HERE, "%s.npz" % name)
    np.savez_compressed(path, nx=NX, seed=SEED, steps=STEPS, **{"out_" + k: v for k, v in out.items() if k in tc.OUT_FIELDS})
    print(path, os.path.getsize(path))

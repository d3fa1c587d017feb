#!/usr/bin/env python
"""Times the thermo() kernel on one GPU and the CPU oracle beside it.

    python profiles/thermo_bench.py [--mesh 3km] [--steps 20] [--warmup 5] [--cpu-elements 200000]

Prints one JSON line: elements/s of k_thermo (CUDA events on the handle's stream, working set >> L2), the HBM roofline
fraction from the algorithmic bytes per element (724 B with the default options), the end-to-end rate of a
host-resident step (forcing upload + thermo + diagnostics download) and the single-thread CPU oracle rate on a sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ALGO_BYTES = 724.0     # 34 doubles read + 3 node ids + 8 B of shared wind, 55 doubles written


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="3km")
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cpu-elements", type=int, default=200000)
    ap.add_argument("--random-regimes", action="store_true", help="ice regime drawn per element (worst case for divergence)")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    from nextsim_b200 import capi, cases, partition as pt, synthetic as syn
    from oracle import thermo as oth

    c = cases.make_case(a.mesh + "_stable", nx=a.nx)
    s = cases.make_solvers(c)[0]
    lm = c.lms[0]
    ne, nn = c.gm.ne, c.gm.nn
    cx, cy = syn.element_centroids(c.gm)
    S = syn.make_thermo_state(ne, nn, seed=11, young=True, centroids=None if a.random_regimes else (cx, cy, c.gm.nx * c.gm.h))
    loc = {k: pt.scatter_elem(lm, S[k]) for k in syn.THERMO_FORCING + syn.THERMO_STATE + syn.THERMO_ICE}
    s.upload(**{k: loc[k] for k in syn.THERMO_ICE}, M_wind=pt.scatter_nodal2(lm, S["M_wind"], nn))
    s.thermo_upload(**{k: loc[k] for k in syn.THERMO_FORCING + syn.THERMO_STATE})
    p = capi.thermo_default_params()
    t0 = 43133.25
    stream = torch.cuda.ExternalStream(capi.lib().nsx_get_stream(s.h), device=torch.device("cuda", 0))
    for k in range(a.warmup):
        s.thermo(p, 200, t0 + k * 200 / 86400.0)
    s.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    with torch.cuda.stream(stream):
        for k, (e0, e1) in enumerate(ev):
            e0.record(stream)
            s.thermo(p, 200, t0 + (a.warmup + k) * 200 / 86400.0)
            e1.record(stream)
    s.synchronize()
    ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
    rate = ne / (np.median(ms) * 1e-3)

    # end to end: this step's forcing from pinned host memory, thermo(), the coupling diagnostics back
    forcing = {k: torch.from_numpy(loc[k]).pin_memory().numpy() for k in syn.THERMO_FORCING}
    diag = ("D_Qa", "D_Qo", "D_delS", "D_fwflux", "D_brine", "D_evap", "D_rain", "D_tau_ow", "D_albedo", "D_vice_melt")
    for _ in range(2):
        s.thermo_upload(**forcing); s.thermo(p, 200, t0); s.thermo_download(*diag)
    w0 = time.perf_counter()
    n_e2e = 5
    for k in range(n_e2e):
        s.thermo_upload(**forcing); s.thermo(p, 200, t0 + k * 200 / 86400.0); s.thermo_download(*diag)
    e2e_s = (time.perf_counter() - w0) / n_e2e

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 7700.0))

    # CPU oracle (single thread) on a sample of the same state
    m = min(a.cpu_elements, ne)
    tri0 = (c.gm.tri - 1)[:m]
    F = {k: S[k][:m] for k in syn.THERMO_FORCING + syn.THERMO_STATE + syn.THERMO_ICE}
    q = oth.default_params()
    oth.thermo(q, 200, t0, tri0, nn, S["M_wind"], S["M_VT"], S["M_ocean"], F)
    w0 = time.perf_counter()
    oth.thermo(q, 200, t0, tri0, nn, S["M_wind"], S["M_VT"], S["M_ocean"], F)
    cpu_s = time.perf_counter() - w0

    out = {"tag": a.tag, "regimes": "random per element" if a.random_regimes else "smooth in space",
           "metric": "thermo_elements_per_second", "value": rate, "unit": "elements/s", "mesh": a.mesh, "elements": ne,
           "ms_per_call": {"median": float(np.median(ms)), "min": float(ms.min()), "max": float(ms.max())},
           "steps": a.steps, "warmup": a.warmup, "dtype": "f64",
           "roofline": {"bound": "hbm", "achieved": rate * ALGO_BYTES / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": rate * ALGO_BYTES / 1e9 / peak, "algorithmic_bytes_per_element": ALGO_BYTES,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "B200_PROFILING.md fallback"},
           "e2e": {"value": ne / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": len(forcing) * ne * 8,
                   "d2h_bytes_per_step": len(diag) * ne * 8},
           "cpu_baseline": {"value": m / cpu_s, "unit": "elements/s", "cores": 1, "kind": "port",
                            "sample": "%d elements of the same state, oracle.thermo (host build of the same element function)" % m}}
    print(json.dumps(out))
    s.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- element-subcycles/s (FP64) of the explicit momentum/rheology hot path on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload 10km|3km|1km|toy] [--dyn bbm|mevp|evp] [--scaling weak|strong] [--no-north-star]

A "step" is one model time step of the path: explicitSolve() (prep, `substeps` sub-cycles, open-water
smoother, tau_w) followed by update().  N=1 runs BASELINE.json configs[1] (synthetic 10 km mesh, 199 712
elements, BBM, 120 sub-cycles).  For N>1 (one process per GPU under torchrun) the mesh is partitioned with
the reference's own node/element ownership rules and ghosts are exchanged over NVLink every sub-cycle;
the headline line is weak scaling (about 2e5 elements per GPU).  In the same process the N>1 runs then
  * check parity of the real NVLink exchange against the oracle (`parity`, exit 1 above 1e-9), and
  * time the north_star multi-GPU configurations (`north_star`): 3 km mesh strong-scaled at N=2,4; at N=8 the
    1 km mesh strong-scaled and the 3 km-mesh-per-GPU weak point.
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads, one partition per
thread) on the same workload; that arm never loads libnsx.so.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {"bbm": 264.0, "evp": 264.0, "mevp": 200.0}     # SURVEY.md 8(d) / BASELINE.md section 3
METRIC = "element-subcycles/sec (FP64)"
UNIT = "element-subcycles/s"
PARITY_TOL = 1e-9

# per-step host<->device traffic of a host that keeps thermodynamics and forcing
E2E_UP = ("M_wind", "M_ocean", "M_ssh", "M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young",
          "M_hs_young", "M_damage", "M_time_relaxation_damage")
E2E_DOWN = ("M_VT", "M_UM", "M_UT", "D_tau_a", "D_tau_w", "M_sigma", "M_damage", "M_conc", "M_thick",
            "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young", "M_thick_myi", "M_conc_myi",
            "M_ridge_ratio", "M_surface")

KERNEL_OF_PATH = {
    "direct": "one sub-cycle = k_element_direct + k_node_direct (working set resident in L2)",
    "tiles": "one sub-cycle = k_subcycle (persistent TMA tile pipeline streaming from HBM)",
    "resident": "k_resident: ONE launch per model step = all sub-cycles + the 50 smoother sweeps, state resident in shared "
                "memory / registers, 16-byte {value, epoch} mailbox stores between tiles and between GPUs",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="10km", choices=["toy", "10km", "3km", "1km"])
    ap.add_argument("--dyn", default="bbm", choices=["bbm", "mevp", "evp"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-north-star", action="store_true", help="N>1: skip the north_star configurations")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the NVLink parity preflight")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--soak-seconds", type=float, default=1.0, help="same step repeated after the timed region for clock samples")
    ap.add_argument("--nx", type=int, default=0, help="override the mesh size (quads per side); experiments only")
    return ap.parse_args()


def workload_nx(workload, scaling, gpus, nx_override=0):
    from nextsim_b200 import synthetic as syn
    nx, h = syn.SIZES[workload]
    if nx_override:
        return nx_override, h
    if scaling == "weak" and gpus > 1:
        nx = int(round(nx * np.sqrt(gpus)))
    return nx, h


def workload_name(workload, dyn, ne):
    return "synthetic %s-class triangular mesh, %d elements, %s, %d sub-cycles/step, explicitSolve+update" % (
        workload, ne, dyn.upper() if dyn != "mevp" else "mEVP", 120)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons every 20 ms while the GPU runs the benchmark's step (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_timed_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, in_timed = [], [], set(), 0
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, r in self.rows:
            if t_begin is not None and ts < t_begin:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            if t_timed_end is not None and ts <= t_timed_end:
                in_timed += 1
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_steps": in_timed,
                "window": "the K timed steps plus the same step repeated for the soak time right after them, sampled every 20 ms"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement on the host cores (reported baseline; also `--impl reference`).
# Nothing here loads libnsx.so: option defaults come from the oracle's own restatement of options.cpp.
# ---------------------------------------------------------------------------------------------------------
def oracle_defaults():
    from nextsim_b200 import capi            # importing the module defines the structs; it does not load the library
    from oracle import oracle as orc
    q, c_lab, alea, trd = orc.default_params()
    p = capi.NsxDynParams()
    for name, _ in orc.OrcParams._fields_:
        if name != "pad_":
            setattr(p, name, getattr(q, name))
    p.use_coriolis = 1
    p.C_lab, p.alea_factor, p.time_relaxation_damage_days = c_lab, alea, trd
    return p


def cpu_arm(args, nx, target_seconds, steps=1, warmup=0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nextsim_b200 import cases
    import oracle_bridge as ob
    from oracle import oracle as orc
    cores = max(1, min(os.cpu_count() or 1, 64))
    c = cases.make_case(args.workload, nranks=cores, dyn=args.dyn, nx=nx, defaults=oracle_defaults())
    ranks = ob.make_ranks(c, fast=True)
    q = ob.orc_params(c.params)
    ne = c.gm.ne
    t_probe = orc.time_subcycles(ranks, q, 2, threads=cores > 1)
    rate = ne * 2 / t_probe
    nsub = int(max(2, min(120, round(target_seconds * rate / ne))))
    times = []
    for i in range(warmup + steps):
        t = orc.time_subcycles(ranks, q, nsub, threads=cores > 1)
        if i >= warmup:
            times.append(t)
    tm = float(np.mean(times))
    return {"value": ne * nsub / tm, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d sub-cycles of the %d-element mesh per step, %d partitions as threads (oracle -O3, "
                      "in-memory updateGhosts); sub-cycle loop only" % (nsub, ne, cores)}, tm, ne, c.gm.nn


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, h = workload_nx(args.workload, args.scaling, args.gpus, args.nx)
    per_step = max(2.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    cb, tm, ne, nn = cpu_arm(args, nx, per_step, steps=args.steps, warmup=args.warmup)
    loaded = any("libnsx" in l for l in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tm * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, args.dyn, ne), "elements": ne, "nodes": nn, "substeps": 120,
                       "parallelism": "cpu threads x%d" % cb["cores"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "product_library_loaded": loaded}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class Env:
    pass


def make_rank_solver(E, c, **options):
    """This rank's handle for case `c`, state uploaded, halo wired over CUDA IPC (one process per GPU)."""
    from nextsim_b200 import capi, cases
    lm, f = c.lms[E.rank], c.local[E.rank]
    S = capi.Solver(lm, device=E.local_rank, **options)
    S.set_params(c.params)
    S.upload(**{k: f[k] for k in cases.UPLOAD_KEYS})
    if E.world > 1:
        blobs = {p: S.halo_blob(p) for p in S.peers}
        allb = [None] * E.world
        E.dist.all_gather_object(allb, blobs)
        for p in S.peers:
            S.halo_connect_blob(p, allb[p][E.rank])
        S.halo_finalize()
        E.dist.barrier()
    return S


def barrier(E):
    E.torch.cuda.synchronize()
    if E.dist is not None:
        E.dist.barrier()
        E.torch.cuda.synchronize()


def timed_steps(E, S, c, steps, warmup, soak_seconds, sample_clocks=True):
    """W untimed steps, then EXACTLY `steps` steps timed with CUDA events on the launching stream (L2 flushed before each
    step), barrier + synchronize on both sides, max over ranks.  The clock sampler also covers `soak_seconds` of the same
    step repeated right after the timed region so that a millisecond-scale region still yields a clock record."""
    torch = E.torch
    stream = torch.cuda.ExternalStream(E.capi.lib().nsx_get_stream(S.h), device=torch.device("cuda", E.local_rank))

    def step():
        S.explicit_solve()
        S.update()

    for _ in range(warmup):
        step()
    barrier(E)
    sampler = ClockSampler(E.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
        time.sleep(0.06)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    sub_ms, prep_ms, ow_ms, upd_ms, launches = [], [], [], [], 0
    barrier(E)
    t_begin = time.perf_counter()
    for i in range(steps):
        with torch.cuda.stream(stream):
            E.flush.zero_()
            ev0[i].record(stream)
            step()
            ev1[i].record(stream)
        t = S.timing()          # syncs the stream; per-phase CUDA-event times of this step
        sub_ms.append(t.subcycle_ms); prep_ms.append(t.prep_ms); ow_ms.append(t.ow_smoother_ms); upd_ms.append(t.update_ms)
        launches += t.n_launches + 1
    barrier(E)
    t_timed_end = time.perf_counter()
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    tot = torch.tensor([sum(step_ms), sum(sub_ms)], dtype=torch.float64, device="cuda")
    if E.dist is not None:
        E.dist.all_reduce(tot, op=E.dist.ReduceOp.MAX)
    total_ms, total_sub_ms = float(tot[0]), float(tot[1])
    clocks = None
    if sampler:
        # every rank runs the SAME number of soak steps (the exchange is collective): sized from the all-reduced step time
        n_soak = int(max(8, min(20000, soak_seconds / max(1e-6, total_ms * 1e-3 / steps))))
        for i in range(n_soak):
            step()
            if i % 16 == 15:
                S.synchronize()
        barrier(E)
        clocks = sampler.stop(t_begin, t_timed_end)
    nsub = c.params.substeps
    ne = c.gm.ne
    return {"value": ne * nsub * steps / (total_ms * 1e-3), "sub_value": ne * nsub * steps / (total_sub_ms * 1e-3),
            "total_ms": total_ms, "ms_per_step": total_ms / steps, "launches": launches, "clocks": clocks,
            "phase_ms": {"prep": float(np.mean(prep_ms)), "subcycles": float(np.mean(sub_ms)),
                         "ow_smoother": float(np.mean(ow_ms)), "update": float(np.mean(upd_ms))},
            "us_per_subcycle_this_rank": float(np.mean(sub_ms)) / nsub * 1e3}


def roofline(dyn, path, n_elements_this_rank, us_per_subcycle, traffic=None):
    peak, peak_src = measured_peak()
    t_sub = us_per_subcycle * 1e-6
    achieved = ALGO_BYTES[dyn] * n_elements_this_rank / t_sub / 1e9
    note = {"tiles": "HBM-bound: the working set streams from HBM every sub-cycle",
            "direct": "NOTIONAL HBM fraction: the working set is resident in the 126 MB L2, DRAM traffic inside the loop is ~0",
            "resident": "NOTIONAL HBM fraction: the sub-cycle state is resident in shared memory / registers, neither HBM nor "
                        "L2 bandwidth bounds the loop (shared-memory wavefronts, issue slots and the FP64 pipe do)"}[path]
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "note": note,
            "algorithmic_bytes_per_launch": ALGO_BYTES[dyn] * n_elements_this_rank * (1 if path != "resident" else 120),
            "launch_unit": "one sub-cycle" if path != "resident" else
                           "one model step = 120 sub-cycles + the smoother in one launch; the CUDA-event time of the launch is split "
                           "into its sub-cycle and smoother shares by %globaltimer stamps taken inside the kernel",
            "kernel": KERNEL_OF_PATH[path] + ", %g B/element-sub-cycle algorithmic" % ALGO_BYTES[dyn],
            "path": path, "us_per_subcycle": us_per_subcycle}


def parity_preflight(E, path):
    """One model step + update() of a small stable case partitioned over the N real GPUs, through the same path and the same
    NVLink exchange as the timed workload, against the oracle's MPI-style replay on rank 0 (FE.cpp:13963-13996)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nextsim_b200 import cases
    nx = 96 if E.world <= 4 else 128
    c = cases.make_case("10km_stable", nranks=E.world, dyn="bbm", nx=nx, open_east=True, only_rank=E.rank)
    S = make_rank_solver(E, c, path=path)
    used = S.path
    S.explicit_solve()
    S.update()
    keys = ("M_VT", "M_UM", "M_UT", "M_sigma", "M_damage", "D_tau_w", "M_conc", "M_thick")
    got = S.download(*keys)
    S.close()
    allg = [None] * E.world
    E.dist.all_gather_object(allg, got)
    res = {"worst_rel_l2": None, "nranks": E.world, "path": used, "tolerance": PARITY_TOL,
           "case": "10km_stable nx=%d (%d elements), BBM, one step + update(), open east boundary" % (nx, c.gm.ne)}
    if E.rank == 0:
        import oracle_bridge as ob
        from oracle import oracle as orc
        cfull = cases.make_case("10km_stable", nranks=E.world, dyn="bbm", nx=nx, open_east=True)
        ranks = ob.make_ranks(cfull, fast=True)
        q = ob.orc_params(cfull.params)
        orc.explicit_solve(ranks, q)
        for R in ranks:
            R.update(q)
        worst = 0.0
        for r, R in enumerate(ranks):
            ref = ob.get_state(R, keys)
            for k in keys:
                pairs = zip(allg[r][k], ref[k]) if k == "M_sigma" else [(allg[r][k], ref[k])]
                for g, h in pairs:
                    worst = max(worst, float(ob.rel_l2(g, h)))
        res["worst_rel_l2"] = worst
    E.dist.barrier()
    return res


def run_ours(args):
    import torch
    from nextsim_b200 import capi, cases
    E = Env()
    E.torch, E.capi = torch, capi
    E.rank = int(os.environ.get("RANK", "0"))
    E.world = int(os.environ.get("WORLD_SIZE", "1"))
    E.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if E.world != args.gpus and E.world == 1 and args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun with %d processes" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(E.local_rank)
    E.dist = None
    if E.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", E.local_rank))
        E.dist = dist
    E.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2
    rank, world = E.rank, E.world

    nx, h = workload_nx(args.workload, args.scaling, world, args.nx)
    c = cases.make_case(args.workload, nranks=world, dyn=args.dyn, nx=nx, only_rank=rank)
    lm, f = c.lms[rank], c.local[rank]
    ne_global = c.gm.ne
    S = make_rank_solver(E, c)
    path = S.path

    # ---- parity of the real multi-GPU exchange, before anything is timed ----
    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_preflight(E, path)
        ok = torch.tensor([1.0 if (rank != 0 or parity["worst_rel_l2"] <= PARITY_TOL) else 0.0], dtype=torch.float64, device="cuda")
        E.dist.all_reduce(ok, op=E.dist.ReduceOp.MIN)
        if float(ok[0]) == 0.0:
            if rank == 0:
                print(json.dumps({"parity": parity, "error": "multi-GPU parity preflight failed"}), flush=True)
            E.dist.barrier()
            sys.exit(1)

    # ---- device-resident timed region ----
    m = timed_steps(E, S, c, args.steps, args.warmup, args.soak_seconds)
    nsub = c.params.substeps

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region ----
    host_up = {k: np.ascontiguousarray(f[k]).copy() for k in E2E_UP}
    host_dn = S.download(*E2E_DOWN)
    L = capi.lib()
    pinned = []
    for a in list(host_up.values()) + [x for k, v in host_dn.items() for x in (v if isinstance(v, list) else [v])]:
        if L.nsx_host_register(a.ctypes.data, a.nbytes) == 0:
            pinned.append(a)
    h2d = sum(a.nbytes for a in host_up.values())
    d2h = sum(x.nbytes for k, v in host_dn.items() for x in (v if isinstance(v, list) else [v]))

    def e2e_step():
        S.upload(**host_up)
        S.explicit_solve()
        S.update()
        S.download(*E2E_DOWN, out=host_dn)

    e2e_step()
    barrier(E)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier(E)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if E.dist is not None:
        E.dist.all_reduce(t_e2e, op=E.dist.ReduceOp.MAX)
    e2e_s = float(t_e2e[0])
    e2e_value = ne_global * nsub * args.steps / e2e_s
    copy_s = max(1e-9, e2e_s - m["total_ms"] * 1e-3)
    for a in pinned:
        L.nsx_host_unregister(a.ctypes.data)

    chk = S.check()
    # ---- SURVEY 8(f) rows 1-2 (device-side regrid check / diagnostics / forcing interpolation): explained numbers,
    # not part of `value`.  GB/s against the algorithmic bytes of each map.
    next_rows = None
    if world == 1:
        reps = 20
        stream = torch.cuda.ExternalStream(L.nsx_get_stream(S.h), device=torch.device("cuda", E.local_rank))
        S.forcing_load("M_wind", 0, f["M_wind"]); S.forcing_load("M_wind", 1, f["M_wind"])

        def timed(fn):
            fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                a.record(stream)
                for _ in range(reps):
                    fn()
                b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps * 1e3
        us_diag = timed(S.update_ice_diagnostics)
        us_forc = timed(lambda: S.forcing_apply("M_wind", True, 0.4, 0.0, 1.0))
        t0 = time.perf_counter()
        for _ in range(reps):
            rg = S.check_regridding(10.0)
        us_regrid = (time.perf_counter() - t0) / reps * 1e6
        nb = {"diag": 156.0 * lm.num_elements, "forcing": 48.0 * lm.num_nodes}
        # row 3: thermo() on the resident state, then FiniteElement::step() between remeshes without the per-step transfers
        # of `e2e` (forcing time slices live on the device, thermo() runs there; the host reads the regrid decision)
        from nextsim_b200 import synthetic as syn, partition as pt
        cx, cy = syn.element_centroids(c.gm)
        TS = syn.make_thermo_state(c.gm.ne, c.gm.nn, seed=11, young=True, centroids=(cx, cy, c.gm.nx * c.gm.h))
        S.thermo_upload(**{k: pt.scatter_elem(lm, TS[k]) for k in syn.THERMO_FORCING + syn.THERMO_STATE})
        tp = capi.thermo_default_params(dtime_step=float(c.params.dtime_step))
        dt_i = int(c.params.dtime_step)
        tnow = [43133.25]
        us_thermo = timed(lambda: S.thermo(tp, dt_i, tnow[0]))
        TH_FORCING = ("M_tair", "M_dair", "M_mslp", "M_Qsw_in", "M_tcc", "M_precip")
        for k in TH_FORCING:
            a = pt.scatter_elem(lm, TS[k])
            S.thermo_forcing_load(k, 0, a); S.thermo_forcing_load(k, 1, a)
        for k in ("M_ocean", "M_ssh"):
            S.forcing_load(k, 0, f[k]); S.forcing_load(k, 1, f[k])
        S.upload(**{k: f[k] for k in cases.UPLOAD_KEYS})          # back to the case's initial state

        def resident_step(k):
            t = 43133.25 + k * dt_i / 86400.0
            for name in ("M_wind", "M_ocean", "M_ssh"):
                S.forcing_apply(name, True, t, 43133.25, 43133.5)
            for name in TH_FORCING:
                S.thermo_forcing_apply(name, True, t, 43133.25, 43133.5)
            S.thermo(tp, dt_i, t)
            S.explicit_solve()
            S.update()
            return S.check_regridding(10.0)                      # synchronises and reads the decision back
        resident_step(0)
        t0 = time.perf_counter()
        for k in range(args.steps):
            resident_step(1 + k)
        rs_s = (time.perf_counter() - t0) / args.steps
        chk_rs = S.check()
        resident = {"value": ne_global * nsub / rs_s, "unit": UNIT, "ms_per_step": rs_s * 1e3,
                    "fraction_of_device_rate": ne_global * nsub / rs_s / m["value"],
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": ctypes.sizeof(capi.NsxRegrid),
                    "per_step": "time interpolation of wind / ocean / ssh and of 6 element forcing fields on the device, thermo(dt), "
                                "explicitSolve(), update(), checkRegridding() with its result read back; forcing time slices were "
                                "loaded once (they change every few hours of model time), the state never leaves the GPU",
                    "check": {"n_nan": chk_rs.n_nan, "n_range": chk_rs.n_range, "max_speed": chk_rs.max_speed}}
        next_rows = {"thermo_us": us_thermo, "thermo_elements_per_s": lm.num_elements / us_thermo * 1e6,
                     "thermo_GBps_of_724B_per_element": 724.0 * lm.num_elements / us_thermo * 1e-3,
                     "thermo_note": "instruction-bound FP64 kernel (divisions, exp / log / cbrt / atan), see profiles/r2_thermo_v4.txt",
                     "resident_step": resident,
                     "update_ice_diagnostics_us": us_diag, "update_ice_diagnostics_GBps": nb["diag"] / us_diag * 1e-3,
                     "forcing_apply_wind_us": us_forc, "forcing_apply_wind_GBps": nb["forcing"] / us_forc * 1e-3,
                     "check_regridding_us_incl_sync_and_readback": us_regrid, "min_angle_deg": rg.min_angle,
                     "launches": 3 * reps + 3 + (reps + 1) + (args.steps + 1) * 10}
    traffic = None
    try:                                  # ncu dram__bytes of one launch; only meaningful for the HBM-streaming tile path
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s/%s/%s" % (args.workload, args.dyn, path)
        if world == 1 and not args.nx and path == "tiles" and key in tj:
            traffic = tj[key]["bytes"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.dyn, ne_global), "elements": ne_global, "nodes": c.gm.nn,
                   "substeps": nsub, "dt_s": c.params.dtime_step,
                   "parallelism": "1 GPU" if world == 1 else
                                  "mesh partitioned over %d GPUs, NVLink ghost push + flags inside the sub-cycle launch" % world,
                   "l2": "flushed (256 MiB write) before every timed step", "path": path},
        "clocks": m["clocks"],
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "fraction_of_device_rate": e2e_value / m["value"],
                "copy_GBps": (h2d + d2h) * args.steps / copy_s / 1e9,
                "note": "upload -> explicitSolve -> update -> download per step through the C ABI; the copies are serial "
                        "with the solve (the host needs the result before it can produce the next input), so the gap to the "
                        "device rate is PCIe time = bytes / copy_GBps"},
        "gpu_launches": m["launches"],
        "roofline": roofline(args.dyn, path, lm.num_elements, m["us_per_subcycle_this_rank"], traffic),
        "subcycle_loop": {"value": m["sub_value"], "unit": UNIT},
        "whole_step_notional_hbm_frac": m["value"] / world * ALGO_BYTES[args.dyn] / (measured_peak()[0] * 1e9),
        "simulated_days_per_wallhour": (args.steps * c.params.dtime_step / 86400.0) / (m["total_ms"] * 1e-3 / 3600.0),
        "phase_ms": m["phase_ms"],
        "check": {"n_nan": chk.n_nan, "n_speed": chk.n_speed, "max_speed": chk.max_speed},
    }
    if parity is not None:
        line["parity"] = parity
    if next_rows is not None:
        line["next_rows"] = next_rows
    S.close()

    # ---- north_star multi-GPU configurations, each with its own clock record ----
    if world > 1 and not args.no_north_star and not args.nx:
        plans = {2: [("3km", "strong")], 4: [("3km", "strong")], 8: [("1km", "strong"), ("3km", "weak")]}.get(world, [])
        ns = {}
        for wl, sc in plans:
            t_setup = time.perf_counter()
            nx2, _ = workload_nx(wl, sc, world)
            c2 = cases.make_case(wl, nranks=world, dyn="bbm", nx=nx2, only_rank=rank)
            S2 = make_rank_solver(E, c2)
            steps2 = max(2, min(args.steps, 5))
            m2 = timed_steps(E, S2, c2, steps2, 3, args.soak_seconds)
            chk2 = S2.check()
            bad = torch.tensor([chk2.n_nan + chk2.n_range], dtype=torch.float64, device="cuda")
            E.dist.all_reduce(bad, op=E.dist.ReduceOp.SUM)
            ns["%s_%s" % (wl, sc)] = {
                "config": {"workload": workload_name(wl, "bbm", c2.gm.ne), "elements": c2.gm.ne, "scaling": sc,
                           "elements_this_rank": c2.lms[rank].num_elements, "path": S2.path},
                "value": m2["value"], "unit": UNIT, "steps": steps2, "warmup": 3, "ms_per_step": m2["ms_per_step"],
                "phase_ms": m2["phase_ms"], "clocks": m2["clocks"],
                "roofline": roofline("bbm", S2.path, c2.lms[rank].num_elements, m2["us_per_subcycle_this_rank"]),
                "check_bad_entries_all_ranks": int(bad[0]), "setup_seconds": None}
            S2.close()
            ns["%s_%s" % (wl, sc)]["setup_seconds"] = time.perf_counter() - t_setup
        line["north_star"] = ns

    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cb, _, _, _ = cpu_arm(args, nx, args.cpu_seconds)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if E.dist is not None:
        E.dist.barrier()
        E.dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

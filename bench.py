#!/usr/bin/env python
"""bench.py -- element-subcycles/s (FP64) of the explicit momentum/rheology hot path on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload 10km|3km|1km|toy] [--dyn bbm|mevp|evp] [--scaling weak|strong]

A "step" is one model time step of the path: explicitSolve() (prep, `substeps` sub-cycles, open-water
smoother, tau_w) followed by update().  N=1 runs BASELINE.json configs[1] (synthetic 10 km mesh, 199 712
elements, BBM, 120 sub-cycles).  For N>1 (one process per GPU under torchrun) the mesh is partitioned with
the reference's own node/element ownership rules and ghosts are exchanged over NVLink every sub-cycle;
default is weak scaling (about 2e5 elements per GPU).  `--impl reference` times the CPU restatement of the
reference (oracle/, all host threads, one partition per thread) on the same workload.
One JSON line on stdout (rank 0).
"""
import argparse
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

# per-step host<->device traffic of a host that keeps thermodynamics and forcing
E2E_UP = ("M_wind", "M_ocean", "M_ssh", "M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young",
          "M_hs_young", "M_damage", "M_time_relaxation_damage")
E2E_DOWN = ("M_VT", "M_UM", "M_UT", "D_tau_a", "D_tau_w", "M_sigma", "M_damage", "M_conc", "M_thick",
            "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young", "M_thick_myi", "M_conc_myi",
            "M_ridge_ratio", "M_surface")


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
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--nx", type=int, default=0, help="override the mesh size (quads per side); experiments only")
    return ap.parse_args()


def workload_nx(args):
    from nextsim_b200 import synthetic as syn
    nx, h = syn.SIZES[args.workload]
    if args.nx:
        return args.nx, h
    if args.scaling == "weak" and args.gpus > 1:
        nx = int(round(nx * np.sqrt(args.gpus)))
    return nx, h


def workload_name(args, ne):
    return "synthetic %s-class triangular mesh, %d elements, %s, %d sub-cycles/step, explicitSolve+update" % (
        args.workload, ne, args.dyn.upper() if args.dyn != "mevp" else "mEVP", 120)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement on the host cores (reported baseline; also `--impl reference`)
# ---------------------------------------------------------------------------------------------------------
def cpu_arm(args, nx, target_seconds, steps=1, warmup=0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nextsim_b200 import cases
    import oracle_bridge as ob
    from oracle import oracle as orc
    cores = max(1, min(os.cpu_count() or 1, 64))
    c = cases.make_case(args.workload if args.workload != "toy" else "toy", nranks=cores, dyn=args.dyn, nx=nx)
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
                      "in-memory updateGhosts); sub-cycle loop only" % (nsub, ne, cores)}, tm, ne


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, h = workload_nx(args)
    per_step = max(2.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    cb, tm, ne = cpu_arm(args, nx, per_step, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tm * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, ne), "parallelism": "cpu threads x%d" % cb["cores"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from nextsim_b200 import capi, cases
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d processes" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nx, h = workload_nx(args)
    c = cases.make_case(args.workload, nranks=world, dyn=args.dyn, nx=nx, only_rank=rank)
    lm, f = c.lms[rank], c.local[rank]
    ne_global = c.gm.ne
    S = capi.Solver(lm, device=local_rank)
    S.set_params(c.params)
    S.upload(**{k: f[k] for k in cases.UPLOAD_KEYS})
    if world > 1:
        blobs = {p: S.halo_blob(p) for p in S.peers}
        allb = [None] * world
        dist.all_gather_object(allb, blobs)
        for p in S.peers:
            S.halo_connect_blob(p, allb[p][rank])
        S.halo_finalize()
        dist.barrier()

    stream = torch.cuda.ExternalStream(capi.lib().nsx_get_stream(S.h), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        S.explicit_solve()
        S.update()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- device-resident timed region: K steps, CUDA events on the launching stream, L2 flushed between ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sub_ms, prep_ms, ow_ms, upd_ms, launches = [], [], [], [], 0
    barrier()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
            ev0[i].record(stream)
            step()
            ev1[i].record(stream)
        t = S.timing()          # syncs the stream; per-phase CUDA-event times of this step
        sub_ms.append(t.subcycle_ms); prep_ms.append(t.prep_ms); ow_ms.append(t.ow_smoother_ms); upd_ms.append(t.update_ms)
        launches += t.n_launches + 1
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    tot_ms = torch.tensor([sum(step_ms), sum(sub_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    total_ms, total_sub_ms = float(tot_ms[0]), float(tot_ms[1])
    nsub = c.params.substeps
    value = ne_global * nsub * args.steps / (total_ms * 1e-3)
    sub_value = ne_global * nsub * args.steps / (total_sub_ms * 1e-3)

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
        step()
        S.download(*E2E_DOWN, out=host_dn)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = ne_global * nsub * args.steps / float(t_e2e[0])
    for a in pinned:
        L.nsx_host_unregister(a.ctypes.data)

    chk = S.check()
    # ---- SURVEY 8(f) rows 1-2 (device-side regrid check / diagnostics / forcing interpolation): explained numbers,
    # not part of `value`.  GB/s against the algorithmic bytes of each map.
    next_rows = None
    if world == 1:
        reps = 20
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
        nb = {"diag": 156.0 * lm.num_elements, "forcing": 48.0 * lm.num_nodes, "regrid": 28.0 * lm.num_elements}
        next_rows = {"update_ice_diagnostics_us": us_diag, "update_ice_diagnostics_GBps": nb["diag"] / us_diag * 1e-3,
                     "forcing_apply_wind_us": us_forc, "forcing_apply_wind_GBps": nb["forcing"] / us_forc * 1e-3,
                     "check_regridding_us_incl_sync_and_readback": us_regrid, "min_angle_deg": rg.min_angle,
                     "launches": 3 * reps + 3}
    path = S.path
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s/%s/%s" % (args.workload, args.dyn, path)
        if world == 1 and not args.nx and key in tj:
            traffic = tj[key]["bytes"]
    except Exception:
        pass
    peak, peak_src = measured_peak()
    # roofline unit: one sub-cycle of THIS rank (its element kernel + node kernel [+ halo]); algorithmic bytes
    # = SURVEY 8(d) per-element figure x the elements this rank updates per sub-cycle
    t_sub = float(np.mean(sub_ms)) / nsub * 1e-3
    achieved = ALGO_BYTES[args.dyn] * lm.num_elements / t_sub / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, ne_global), "elements": ne_global, "nodes": c.gm.nn,
                   "substeps": nsub, "dt_s": c.params.dtime_step,
                   "parallelism": "1 GPU" if world == 1 else "mesh partitioned over %d GPUs, NVLink halo push per sub-cycle" % world,
                   "l2": "flushed (256 MiB write) between timed steps"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALGO_BYTES[args.dyn] * lm.num_elements,
                     "kernel": {"direct": "one sub-cycle = k_element_direct + k_node_direct (L2-resident mesh)",
                                "tiles": "one sub-cycle = k_subcycle (TMA tile pipeline)",
                                "resident": "k_resident: ONE launch = all sub-cycles + the 50 smoother sweeps, state resident in "
                                            "shared memory / registers; duration / sub-cycles (smoother time included)"}[path] +
                               ", %g B/element algorithmic" % ALGO_BYTES[args.dyn],
                     "path": path,
                     "us_per_subcycle": t_sub * 1e6},
        "subcycle_loop": {"value": sub_value, "unit": UNIT},
        "simulated_days_per_wallhour": (args.steps * c.params.dtime_step / 86400.0) / (total_ms * 1e-3 / 3600.0),
        "phase_ms": {"prep": float(np.mean(prep_ms)), "subcycles": float(np.mean(sub_ms)),
                     "ow_smoother": float(np.mean(ow_ms)), "update": float(np.mean(upd_ms))},
        "check": {"n_nan": chk.n_nan, "n_speed": chk.n_speed, "max_speed": chk.max_speed},
    }
    if next_rows is not None:
        line["next_rows"] = next_rows
    S.close()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cb, _, _ = cpu_arm(args, nx, args.cpu_seconds)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

"""Multi-GPU parity (one process per GPU, NVLink halo push) against the CPU oracle's MPI-style replay.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/run_multigpu_parity.py [--nx 128] [--dyn bbm] [--steps 2]

Every rank builds its own partition of the same synthetic mesh, wires the halo through the C ABI
(nsx_halo_blob / nsx_halo_connect_blob), runs `steps` x (explicitSolve + update) and sends its owned results
to rank 0, which runs the oracle on all partitions (in-process updateGhosts) and compares: rel-L2 <= 1e-9.
Not collected by pytest (needs torchrun and N GPUs); run with gpurun --gpus N.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=128)
    ap.add_argument("--dyn", default="bbm")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--case", default="10km_stable")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    from nextsim_b200 import capi, cases
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lrank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))

    c = cases.make_case(a.case, nranks=world, dyn=a.dyn, nx=a.nx, open_east=True, only_rank=rank)
    lm, f = c.lms[rank], c.local[rank]
    S = capi.Solver(lm, device=lrank)
    S.set_params(c.params)
    S.upload(**{k: f[k] for k in cases.UPLOAD_KEYS})
    blobs = {p: S.halo_blob(p) for p in S.peers}
    allb = [None] * world
    dist.all_gather_object(allb, blobs)
    for p in S.peers:
        S.halo_connect_blob(p, allb[p][rank])
    S.halo_finalize()
    dist.barrier()
    for _ in range(a.steps):
        S.explicit_solve()
        S.update()
    keys = ("M_VT", "M_UM", "M_UT", "M_sigma", "M_damage", "D_tau_w", "M_conc", "M_thick")
    got = S.download(*keys)
    chk = S.check()
    allg = [None] * world
    dist.all_gather_object(allg, got)
    ok = True
    if rank == 0:
        import oracle_bridge as ob
        from oracle import oracle as orc
        cfull = cases.make_case(a.case, nranks=world, dyn=a.dyn, nx=a.nx, open_east=True)
        ranks = ob.make_ranks(cfull)
        q = ob.orc_params(cfull.params)
        for _ in range(a.steps):
            orc.explicit_solve(ranks, q)
            for R in ranks:
                R.update(q)
        worst = 0.0
        for r, R in enumerate(ranks):
            ref = ob.get_state(R, keys)
            for k in keys:
                pairs = zip(allg[r][k], ref[k]) if k == "M_sigma" else [(allg[r][k], ref[k])]
                for g, h in pairs:
                    e = ob.rel_l2(g, h)
                    worst = max(worst, e)
                    if e > 1e-9:
                        ok = False
                        print("MISMATCH rank %d %s rel-L2 %.3e" % (r, k, e))
        print("multi-GPU parity: %d ranks, %d elements, %s, %d steps: worst rel-L2 %.3e -> %s"
              % (world, cfull.gm.ne, a.dyn, a.steps, worst, "OK" if ok else "FAIL"), flush=True)
    S.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()

"""Latency of one ghost exchange (k_halo_exchange: NVLink push + epoch flag + wait) between real GPUs.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 tests/bench_halo.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from nextsim_b200 import capi, cases
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lrank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    c = cases.make_case("10km_stable", nranks=world, nx=int(316 * world ** 0.5), only_rank=rank)
    S = capi.Solver(c.lms[rank], device=lrank)
    S.set_params(c.params)
    S.upload(**{k: c.local[rank][k] for k in cases.UPLOAD_KEYS})
    blobs = {p: S.halo_blob(p) for p in S.peers}
    allb = [None] * world
    dist.all_gather_object(allb, blobs)
    for p in S.peers:
        S.halo_connect_blob(p, allb[p][rank])
    S.halo_finalize()
    dist.barrier()
    for n in (200, 2000):
        S.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            S.update_ghosts()
        S.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print("ghost exchange: %d ranks, %d sent nodes on rank 0, %.2f us per exchange (launch included, %d calls)"
                  % (world, sum(v.size for v in c.lms[0].send_to.values()), dt / n * 1e6, n), flush=True)
    S.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

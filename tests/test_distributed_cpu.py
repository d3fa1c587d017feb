"""N>1 host logic on CPU: two processes over gloo (127.0.0.1) build their own rank's mesh, exchange halo
descriptors the way bench.py does for the NVLink blobs, and replay updateGhosts with the product's
send/recv lists; the result is checked against the oracle's in-process updateGhosts."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from nextsim_b200 import cases, partition as pt
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = cases.make_case("10km_stable", nranks=world, dyn="bbm", nx=20, open_east=True, only_rank=rank)
    lm, f = c.lms[rank], c.local[rank]
    assert all(c.local[r] is None for r in range(world) if r != rank)
    vt = f["M_VT"].copy()
    rng = np.random.default_rng(100 + rank)
    nn, nd = lm.num_nodes, lm.local_ndof
    vt[:nd] = rng.standard_normal(nd)               # new owned values, ghosts stale
    vt[nn:nn + nd] = rng.standard_normal(nd)
    # descriptor exchange (same shape as the halo blobs: what I expect from each owner)
    mine = {p: lm.recv_from[p].size for p in lm.recv_from}
    allb = [None] * world
    dist.all_gather_object(allb, mine)
    for p, idx in lm.send_to.items():
        assert allb[p][rank] == idx.size, "holder and owner disagree on the list length"
    # updateGhosts: pack [u.. | v..] per holder, exchange, unpack (FE.cpp:13963-13996)
    out = {p: np.concatenate([vt[idx], vt[idx + nn]]) for p, idx in lm.send_to.items()}
    allm = [None] * world
    dist.all_gather_object(allm, out)
    for p, idx in lm.recv_from.items():
        m = allm[p][rank]
        vt[idx] = m[:idx.size]
        vt[idx + nn] = m[idx.size:]
    owned = np.concatenate([vt[:nd], vt[nn:nn + nd]])
    q.put((rank, vt, owned))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_ranks_update_ghosts():
    import torch.multiprocessing as mp
    from nextsim_b200 import cases, partition as pt
    import oracle_bridge as ob
    from oracle import oracle as orc
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, vt, owned = q.get(timeout=120)
        res[r] = (vt, owned)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # oracle: same owned values, its own lists
    c = cases.make_case("10km_stable", nranks=world, dyn="bbm", nx=20, open_east=True)
    ranks = ob.make_ranks(c)
    for r, R in enumerate(ranks):
        lm = c.lms[r]
        v = R.get("M_VT")
        nn, nd = lm.num_nodes, lm.local_ndof
        v[:nd] = res[r][1][:nd]
        v[nn:nn + nd] = res[r][1][nd:]
        R.set("M_VT", v)
    orc.update_ghosts(ranks)
    for r, R in enumerate(ranks):
        assert np.array_equal(R.get("M_VT"), res[r][0]), "rank %d ghosts differ from the oracle" % r

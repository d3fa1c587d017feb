"""State-resident persistent solver (k_resident, NsxCreateOptions.path = RESIDENT; AUTO selects it whenever the mesh fits):
one launch per model step, tile-to-tile synchronisation by release/acquire flags.

Same bar as the other paths: rel-L2 <= 1e-9 against the oracle after a full model step + update(); it also has to agree
with the direct path to rounding.  The in-process groups below run one persistent launch PER RANK concurrently on one
GPU (each on its share of the SMs), so the rank-to-rank flag protocol -- the same code that runs over NVLink between
GPUs -- is exercised by the single-GPU test run."""
import os

import numpy as np
import pytest

from nextsim_b200 import cases
import oracle_bridge as ob
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
KEYS = cases.STATE_OUT


def run_path(c, path, monkeypatch, update=True):
    monkeypatch.setenv("NSX_PATH", path)
    (s,) = cases.make_solvers(c)
    s.explicit_solve()
    out = s.download(*KEYS)
    if update:
        s.update()
        out["update"] = s.download(*cases.UPDATE_OUT)      # M_sigma is rescaled by update(): keep both
    s.close()
    return out


@pytest.mark.parametrize("name,nx,dyn,nsub", [("toy", None, "bbm", 1), ("toy", None, "bbm", 0), ("toy", None, "mevp", 0),
                                              ("10km_stable", 64, "bbm", 0), ("10km_stable", 64, "evp", 0),
                                              ("10km_stable", None, "bbm", 0)])
def test_resident_path(monkeypatch, name, nx, dyn, nsub):
    c = cases.make_case(name, nranks=1, dyn=dyn, nx=nx, open_east=True)
    if nsub:
        c.params.stop_after_substeps = nsub
        c.params.skip_ow_smoother = 1
    got = run_path(c, "resident", monkeypatch, update=not nsub)
    ref_gpu = run_path(c, "direct", monkeypatch, update=not nsub)
    (R,) = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    orc.explicit_solve([R], q)
    ref = ob.get_state(R, KEYS)
    for k in KEYS:
        pairs = zip(got[k], ref[k], ref_gpu[k]) if k == "M_sigma" else [(got[k], ref[k], ref_gpu[k])]
        for g, o, d in pairs:
            assert ob.rel_l2(g, o) <= 1e-9, (k, "vs oracle", ob.rel_l2(g, o))
            assert ob.rel_l2(g, d) <= 1e-11, (k, "vs direct path", ob.rel_l2(g, d))
    if not nsub:
        R.update(q)
        for k in cases.UPDATE_OUT:
            o = ob.get_state(R, (k,))[k]
            gk = got["update"][k]
            pairs = zip(gk, o) if k == "M_sigma" else [(gk, o)]
            for g, r in pairs:
                assert ob.rel_l2(g, r) <= 1e-9, ("update", k)


def test_auto_selects_resident_for_the_headline_mesh(monkeypatch):
    monkeypatch.delenv("NSX_PATH", raising=False)
    c = cases.make_case("10km_stable", nranks=1, dyn="bbm")
    (s,) = cases.make_solvers(c)
    assert s.path == "resident", s.tile_info()
    s.close()


@pytest.mark.parametrize("name,nx,nranks,dyn,open_east", [
    ("toy", None, 2, "bbm", True), ("toy", None, 3, "mevp", False), ("10km_stable", 96, 4, "bbm", True),
    ("10km_stable", 64, 8, "evp", True), ("10km_stable", 160, 2, "bbm", False)])
def test_resident_ranks_exchange_through_flags(monkeypatch, name, nx, nranks, dyn, open_east):
    """updateGhosts() inside the persistent launch: export nodes pushed into the holder's VT buffer, per-link arrival
    counters, release of the exchange epoch, bounded acquire spins (FE.cpp:13963-13996)."""
    from nextsim_b200 import capi
    monkeypatch.setenv("NSX_PATH", "resident")
    c = cases.make_case(name, nranks=nranks, dyn=dyn, nx=nx, open_east=open_east)
    ranks = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    orc.explicit_solve(ranks, q)
    solvers = cases.make_solvers(c)
    assert all(s.path == "resident" for s in solvers)
    capi.group_explicit_solve(solvers)
    for R, s in zip(ranks, solvers):
        got = s.download(*KEYS)
        ref = ob.get_state(R, KEYS)
        for k in KEYS:
            pairs = zip(got[k], ref[k]) if k == "M_sigma" else [(got[k], ref[k])]
            for g, o in pairs:
                assert ob.rel_l2(g, o) <= 1e-9, (k, ob.rel_l2(g, o))
    for R, s in zip(ranks, solvers):
        R.update(q)
        s.update()
        got = s.download(*cases.UPDATE_OUT)
        ref = ob.get_state(R, cases.UPDATE_OUT)
        for k in cases.UPDATE_OUT:
            pairs = zip(got[k], ref[k]) if k == "M_sigma" else [(got[k], ref[k])]
            for g, o in pairs:
                assert ob.rel_l2(g, o) <= 1e-9, ("update", k)
    for s in solvers:
        s.close()

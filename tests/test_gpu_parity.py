"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): relative L2 <= 1e-9 on velocity, stress and damage after one model
time step (120 sub-cycles); we hold every other output of explicitSolve()/update() to the same bound.
"""
import numpy as np
import pytest

from nextsim_b200 import cases
import oracle_bridge as ob
from oracle import oracle as orc

TOL = 1e-9
pytestmark = pytest.mark.gpu


def run_both(c, do_update=True):
    ranks = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    orc.explicit_solve(ranks, q)
    solvers = cases.make_solvers(c)
    if len(solvers) == 1:
        solvers[0].explicit_solve()
    else:
        from nextsim_b200 import capi
        capi.group_explicit_solve(solvers)
    res = {}
    ref = [ob.get_state(R, cases.STATE_OUT) for R in ranks]
    got = [s.download(*cases.STATE_OUT) for s in solvers]
    for k in cases.STATE_OUT:
        res[k] = (got, ref)
        compare(c, k, got, ref)
    if do_update:
        for R in ranks:
            R.update(q)
        for s in solvers:
            s.update()
        ref = [ob.get_state(R, cases.UPDATE_OUT) for R in ranks]
        got = [s.download(*cases.UPDATE_OUT) for s in solvers]
        for k in cases.UPDATE_OUT:
            compare(c, k, got, ref, tag="update:")
    chk = solvers[0].check()
    assert chk.n_nan == 0
    for s in solvers:
        s.close()


def compare(c, k, got, ref, tag=""):
    for r, (g, f) in enumerate(zip(got, ref)):
        if k == "M_sigma":
            for i in range(3):
                e = ob.rel_l2(g[k][i], f[k][i])
                assert e <= TOL, "%s%s[%d] rank %d rel-L2 %.3e" % (tag, k, i, r, e)
        else:
            e = ob.rel_l2(g[k], f[k])
            assert e <= TOL, "%s%s rank %d rel-L2 %.3e" % (tag, k, r, e)


@pytest.mark.parametrize("dyn", ["bbm", "mevp", "evp"])
def test_toy_one_rank(dyn):
    run_both(cases.make_case("toy", nranks=1, dyn=dyn))


@pytest.mark.parametrize("dyn", ["bbm", "mevp"])
def test_toy_one_substep(dyn):
    c = cases.make_case("toy", nranks=1, dyn=dyn)
    c.params.stop_after_substeps = 1
    c.params.skip_ow_smoother = 1
    run_both(c, do_update=False)


@pytest.mark.parametrize("nranks,dyn,open_east", [(2, "bbm", True), (3, "mevp", False), (4, "bbm", False), (8, "evp", True)])
def test_toy_partitioned_group(nranks, dyn, open_east):
    run_both(cases.make_case("toy", nranks=nranks, dyn=dyn, open_east=open_east))


@pytest.mark.parametrize("dyn,young", [("bbm", True), ("bbm", False), ("mevp", True)])
def test_large_state_small_mesh(dyn, young):
    run_both(cases.make_case("10km", nranks=1, dyn=dyn, nx=64, open_east=True, young=young))


def test_10km_full_size_bbm():
    """BASELINE config #2 at full size (199 712 elements): oracle takes a few seconds."""
    run_both(cases.make_case("10km", nranks=1, dyn="bbm"))


def test_multistep_state_stays_resident():
    """Three model steps without re-uploading: device state carries over exactly like the host members."""
    c = cases.make_case("10km", nranks=1, dyn="bbm", nx=48)
    ranks = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    solvers = cases.make_solvers(c)
    for _ in range(3):
        orc.explicit_solve(ranks, q)
        ranks[0].update(q)
        solvers[0].explicit_solve()
        solvers[0].update()
    ref = [ob.get_state(ranks[0], cases.STATE_OUT)]
    got = [solvers[0].download(*cases.STATE_OUT)]
    for k in ("M_VT", "M_sigma", "M_damage", "M_UM"):
        compare(c, k, got, ref, tag="3 steps:")
    solvers[0].close()

"""Pins the connectivity-table part of the oracle (and of the product's host library) against the REFERENCE'S OWN
CODE: contrib/bamg compiled unmodified from /root/reference into oracle/_ref/libref_bamg.so (oracle/ref_bamg/Makefile)
and called the way FiniteElement::distributedMeshProcessing calls it (FE.cpp:77-80, BamgConvertMeshx).  The hot
path reads bamgmesh->NodalElementConnectivity (atmospheric drag loop, FE.cpp:10374-10394) and
bamgmesh->NodalConnectivity (open-water smoother, FE.cpp:10578-10611), whose row ORDER fixes the floating-point
summation order."""
import numpy as np
import pytest

from nextsim_b200 import capi, partition as pt, synthetic as syn
from oracle import oracle as orc
from oracle import ref_bamg
from test_partition_cpu import tags

pytestmark = pytest.mark.skipif(not ref_bamg.available(), reason="oracle/_ref/libref_bamg.so not built and no /root/reference")


def oracle_tables(x, y, tri1):
    R = orc.single_rank_mesh(x, y, tri1)
    R.bamg_tables()
    s = R.sizes()
    return (R.get("NodalElementConnectivity").reshape(x.size, s["nec_width"]),
            R.get("NodalConnectivity").reshape(x.size, s["nc_width"]))


def check_all(x, y, tri1):
    nec_ref, nc_ref, _ = ref_bamg.convert(x, y, tri1)
    nec_o, nc_o = oracle_tables(x, y, tri1)
    assert np.array_equal(nec_ref, nec_o, equal_nan=True), "oracle NodalElementConnectivity != reference bamg"
    assert np.array_equal(nc_ref, nc_o), "oracle NodalConnectivity != reference bamg"
    nec_p, nc_p = pt.bamg_tables(np.asarray(tri1).reshape(-1, 3), x.size)
    assert np.array_equal(nec_ref, nec_p, equal_nan=True) and np.array_equal(nc_ref, nc_p)
    return nec_ref, nc_ref


@pytest.mark.parametrize("nx,open_east,seed", [(4, False, 1), (6, True, 2), (17, False, 3), (32, True, 4), (64, False, 5)])
def test_root_mesh_tables_match_reference_bamg(nx, open_east, seed):
    m = syn.make_mesh(nx, 10e3, seed=seed, open_east=open_east)
    nec, nc = check_all(m.x, m.y, m.tri)
    pm = capi.PartMesh.build(m.x, m.y, m.tri)
    lm = pm.to_local_mesh()
    assert np.array_equal(nec, lm.nodal_element_connectivity, equal_nan=True)
    assert np.array_equal(nc, lm.nodal_connectivity)


@pytest.mark.parametrize("nx,P,method", [(24, 3, "rcb"), (40, 7, "scattered"), (32, 8, "rcb")])
def test_partition_local_mesh_tables_match_reference_bamg(nx, P, method):
    """The local meshes (owned + ghost elements, local numbering) are what the reference hands to BamgConvertMeshx."""
    m = syn.make_mesh(nx, 10e3, open_east=True)
    ep, gp, gv = tags(m, P, method)
    for r in range(P):
        pm = capi.PartMesh.build(m.x, m.y, m.tri, r, P, ep, gp, gv)
        lm = pm.to_local_mesh()
        nec, nc = check_all(lm.x, lm.y, lm.indices)
        assert np.array_equal(nec, lm.nodal_element_connectivity, equal_nan=True)
        assert np.array_equal(nc, lm.nodal_connectivity)


def test_full_size_10km_mesh():
    m = syn.named_mesh("10km")
    nec_ref, nc_ref, _ = ref_bamg.convert(m.x, m.y, m.tri)
    pm = capi.PartMesh.build(m.x, m.y, m.tri)
    lm = pm.to_local_mesh()
    assert np.array_equal(nec_ref, lm.nodal_element_connectivity, equal_nan=True)
    assert np.array_equal(nc_ref, lm.nodal_connectivity)

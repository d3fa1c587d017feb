"""Partition indexing is bit-exact against the oracle's literal replay of GmshMesh::nodalGrid()
(core/src/gmshmesh.cpp:856-1498), FiniteElement::initUpdateGhosts() (FE.cpp:14003-14088),
bcMarkedNodes() (FE.cpp:150-271) and bamg's connectivity tables (contrib/bamg/src/Mesh.cpp:514-865)."""
import numpy as np
import pytest

from nextsim_b200 import partition as pt, synthetic as syn
from oracle import oracle as orc


def tags(m, P, method="rcb"):
    if method == "rcb":
        ep = pt.partition_elements(m.x, m.y, m.tri, P)
    elif method == "strips":
        cx = m.x[m.tri - 1].mean(1)
        ep = np.minimum((cx / (m.nx * m.h) * P).astype(np.int32), P - 1)
    else:                       # scattered: checkerboard of small blocks, many neighbours per rank
        cx = m.x[m.tri - 1].mean(1)
        cy = m.y[m.tri - 1].mean(1)
        b = 4 * m.h
        ep = ((np.floor(cx / b) + 3 * np.floor(cy / b)) % P).astype(np.int32)
    gp, gv = pt.ghost_tags(m.tri, ep, P)
    return ep, gp, gv


@pytest.mark.parametrize("nx,P,method,open_east", [
    (32, 2, "rcb", True), (32, 3, "rcb", False), (32, 4, "rcb", True), (32, 8, "rcb", False),
    (17, 5, "strips", True), (40, 7, "scattered", True), (9, 2, "rcb", False), (64, 16, "rcb", True)])
def test_nodal_grid_bit_exact(nx, P, method, open_east):
    m = syn.make_mesh(nx, 10e3, open_east=open_east)
    ep, gp, gv = tags(m, P, method)
    lms = pt.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv)
    ors = orc.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv)
    owned_nodes = np.concatenate([lm.node_gid[:lm.local_ndof] for lm in lms])
    assert np.array_equal(np.sort(owned_nodes), np.arange(1, m.nn + 1)), "every node owned exactly once"
    owned_el = np.concatenate([lm.elem_gid[:lm.local_nelements] for lm in lms])
    assert np.array_equal(np.sort(owned_el), np.arange(1, m.ne + 1)), "every element owned exactly once"
    for r in range(P):
        lm, R = lms[r], ors[r]
        s = R.sizes()
        assert (s["num_nodes"], s["local_ndof"], s["num_elements"], s["local_nelements"]) == \
            (lm.num_nodes, lm.local_ndof, lm.num_elements, lm.local_nelements)
        assert np.array_equal(R.get("indices").reshape(-1, 3), lm.indices)
        assert np.array_equal(R.get("ghostNodes").reshape(-1, 3), lm.ghostNodes)
        assert np.array_equal(R.get("local_dof_with_ghost_init"), lm.node_gid)
        assert np.array_equal(R.get("local_dof_with_ghost")[:lm.num_nodes], lm.node_rid)
        assert np.array_equal(R.get("triangles_id_with_ghost"), lm.elem_gid)
        assert np.array_equal(R.get("element_partition"), lm.elem_part)
        assert np.array_equal(R.get("local_ghost"), lm.local_ghost)
        assert np.array_equal(R.get("coordX"), lm.x) and np.array_equal(R.get("coordY"), lm.y)
        for p in range(P):
            assert np.array_equal(R.halo(1, p), lm.recv_from.get(p, np.zeros(0, np.int32))), ("recv", r, p)
            assert np.array_equal(R.halo(0, p), lm.send_to.get(p, np.zeros(0, np.int32))), ("send", r, p)
        # lowest rank owns interface nodes; ghost elements only from higher partitions
        assert (lm.elem_part[lm.local_nelements:] > r).all()
        R.bamg_tables()
        nec, nc = pt.bamg_tables(lm.indices, lm.num_nodes)
        s = R.sizes()
        a = R.get("NodalElementConnectivity").reshape(lm.num_nodes, s["nec_width"])
        assert np.array_equal(np.isnan(a), np.isnan(nec))
        assert np.array_equal(np.nan_to_num(a, nan=-1.0), np.nan_to_num(nec, nan=-1.0))
        assert np.array_equal(R.get("NodalConnectivity").reshape(lm.num_nodes, s["nc_width"]), nc)
        R.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
        pt.bc_marked_nodes(lm, m.dirichlet_flags_root, m.neumann_flags_root)
        assert np.array_equal(R.get("M_mask_dirichlet"), lm.mask_dirichlet)
        assert np.array_equal(R.get("M_neumann_flags"), lm.neumann_flags)
        assert np.array_equal(R.get("M_neumann_nodes"), lm.neumann_nodes)
        assert lm.mask_dirichlet[lm.local_ndof:].sum() == 0, "Dirichlet mask covers owned nodes only (Q8)"


def test_bamg_table_ordering():
    """node->element lists in DESCENDING element id, node->node lists in reverse edge-insertion order (Q7)."""
    m = syn.make_mesh(6, 1.0)
    nec, nc = pt.bamg_tables(m.tri, m.nn)
    for n in range(m.nn):
        row = nec[n][~np.isnan(nec[n])]
        assert np.all(np.diff(row) < 0)
        inc = np.nonzero((m.tri - 1 == n).any(1))[0] + 1
        assert np.array_equal(np.sort(row), inc)
        cnt = int(nc[n, -1])
        neigh = nc[n, :cnt].astype(int) - 1
        t = m.tri[(m.tri - 1 == n).any(1)] - 1
        assert set(neigh) == set(np.unique(t)) - {n}
        assert len(set(neigh)) == cnt


def test_update_ghosts_moves_owner_values():
    """FiniteElement::updateGhosts: after the exchange every ghost holds its owner's (u,v)."""
    m = syn.make_mesh(24, 10e3)
    P = 4
    ep, gp, gv = tags(m, P)
    lms = pt.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv)
    ors = orc.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv)
    rng = np.random.default_rng(0)
    glob = rng.standard_normal(2 * m.nn)
    for lm, R in zip(lms, ors):
        v = pt.scatter_nodal2(lm, glob, m.nn)
        v[lm.local_ndof:lm.num_nodes] = np.nan                    # poison the ghosts
        v[lm.num_nodes + lm.local_ndof:] = np.nan
        R.set("M_VT", v)
    orc.update_ghosts(ors)
    for lm, R in zip(lms, ors):
        assert np.array_equal(R.get("M_VT"), pt.scatter_nodal2(lm, glob, m.nn))


def test_single_rank_is_identity():
    m = syn.make_mesh(8, 1.0)
    lm = pt.nodal_grid(1, m.x, m.y, m.tri)[0]
    assert lm.num_nodes == lm.local_ndof == m.nn and lm.num_elements == m.ne
    assert np.array_equal(lm.indices, m.tri) and lm.ghostNodes.sum() == 0


def test_mesh_sizes_match_baseline():
    """toy 2048/1089, 10 km 199712/100489 elements/nodes; all triangles CCW."""
    assert (syn.SIZES["toy"][0] ** 2 * 2, (syn.SIZES["toy"][0] + 1) ** 2) == (2048, 1089)
    assert (syn.SIZES["10km"][0] ** 2 * 2, (syn.SIZES["10km"][0] + 1) ** 2) == (199712, 100489)
    assert (syn.SIZES["3km"][0] ** 2 * 2, (syn.SIZES["3km"][0] + 1) ** 2) == (2000000, 1002001)
    assert (syn.SIZES["1km"][0] ** 2 * 2, (syn.SIZES["1km"][0] + 1) ** 2) == (19996488, 10004569)
    m = syn.named_mesh("toy")
    t = m.tri - 1
    jac = (m.x[t[:, 1]] - m.x[t[:, 0]]) * (m.y[t[:, 2]] - m.y[t[:, 0]]) - \
          (m.x[t[:, 2]] - m.x[t[:, 0]]) * (m.y[t[:, 1]] - m.y[t[:, 0]])
    assert (jac > 0).all()

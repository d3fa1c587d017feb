"""the remesh-time host library (nsx_partmesh_*, no GPU) is bit-exact against the
oracle's literal replay of GmshMesh::nodalGrid() / initUpdateGhosts() / bcMarkedNodes() / bamg's tables, reads the
partitioned msh-2.2 file in both formats the reference reads (core/src/gmshmesh.cpp:133-712), and is much faster
than the std::map-based replay."""
import struct
import time

import numpy as np
import pytest

from nextsim_b200 import capi, partition as pt, synthetic as syn
from oracle import oracle as orc
from test_partition_cpu import tags


def assert_matches_oracle(lm, R, m, P):
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
        assert np.array_equal(R.halo(1, p), lm.recv_from.get(p, np.zeros(0, np.int32))), ("recv", p)
        assert np.array_equal(R.halo(0, p), lm.send_to.get(p, np.zeros(0, np.int32))), ("send", p)
    R.bamg_tables()
    s = R.sizes()
    a = R.get("NodalElementConnectivity").reshape(lm.num_nodes, s["nec_width"])
    assert np.array_equal(a, lm.nodal_element_connectivity, equal_nan=True)
    assert np.array_equal(R.get("NodalConnectivity").reshape(lm.num_nodes, s["nc_width"]), lm.nodal_connectivity)
    R.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
    assert np.array_equal(R.get("M_mask_dirichlet"), lm.mask_dirichlet)
    assert np.array_equal(R.get("M_neumann_flags"), lm.neumann_flags)


@pytest.mark.parametrize("nx,P,method,open_east", [
    (32, 2, "rcb", True), (32, 3, "rcb", False), (32, 8, "rcb", False), (17, 5, "strips", True),
    (40, 7, "scattered", True), (9, 2, "rcb", False), (64, 16, "rcb", True)])
def test_library_bit_exact_against_oracle(nx, P, method, open_east):
    m = syn.make_mesh(nx, 10e3, open_east=open_east)
    ep, gp, gv = tags(m, P, method)
    ors = orc.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv)
    for r in range(P):
        pm = capi.PartMesh.build(m.x, m.y, m.tri, r, P, ep, gp, gv)
        pm.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
        assert_matches_oracle(pm.to_local_mesh(), ors[r], m, P)
        pm.close()


@pytest.mark.parametrize("nx,P,seed", [(12, 3, 0), (20, 6, 1), (16, 11, 2)])
def test_library_bit_exact_on_random_partitions(nx, P, seed):
    """Salt-and-pepper tags (every element on a random rank): many neighbours per rank, most nodes on interfaces."""
    m = syn.make_mesh(nx, 10e3, open_east=True)
    ep = np.random.default_rng(seed).integers(0, P, m.ne).astype(np.int32)
    gp, gv = pt.ghost_tags(m.tri, ep, P)
    ors = orc.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv)
    for r in range(P):
        pm = capi.PartMesh.build(m.x, m.y, m.tri, r, P, ep, gp, gv)
        pm.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
        assert_matches_oracle(pm.to_local_mesh(), ors[r], m, P)


def test_single_rank_library():
    m = syn.make_mesh(12, 1.0)
    pm = capi.PartMesh.build(m.x, m.y, m.tri)
    pm.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
    lm = pm.to_local_mesh()
    assert lm.num_nodes == lm.local_ndof == m.nn and np.array_equal(lm.indices, m.tri) and not lm.send_to
    R = orc.single_rank_mesh(m.x, m.y, m.tri)
    R.bamg_tables()
    s = R.sizes()
    assert np.array_equal(R.get("NodalConnectivity").reshape(m.nn, s["nc_width"]), lm.nodal_connectivity)


# ---- msh 2.2 writer (test side): what Gmsh's GModel::writeMSH emits for a partitioned 2-D mesh
def write_msh(path, m, ep, gp, gv, fmt, edges_first=0, swap=False, bamg=False, blocks=1):
    ne, nn = m.ne, m.nn
    tri = m.tri.copy()
    if bamg:                      # the file holds the vertices so that next_permutation(idx[1:]) restores m.tri
        tri = tri[:, [0, 2, 1]]
    e = ">" if swap else "<"
    with open(path, "wb") as f:
        f.write(b"$MeshFormat\n2.2 %d 8\n" % (1 if fmt == "binary" else 0))
        if fmt == "binary":
            f.write(struct.pack(e + "i", 1) + b"\n")
        f.write(b"$EndMeshFormat\n")
        if edges_first:
            f.write(b'$PhysicalNames\n2\n1 1 "coast"\n1 2 "open"\n$EndPhysicalNames\n')
        f.write(b"$Nodes\n%d\n" % nn)
        perm = np.random.default_rng(1).permutation(nn)       # ids need not be in order
        if fmt == "binary":
            for i in perm:
                f.write(struct.pack(e + "i3d", i + 1, m.x[i], m.y[i], 0.0))
            f.write(b"\n")
        else:
            for i in perm:
                f.write(b"%d %s %s 0\n" % (i + 1, repr(float(m.x[i])).encode(), repr(float(m.y[i])).encode()))
        f.write(b"$EndNodes\n$Elements\n%d\n" % (ne + edges_first))
        num = 1
        for k in range(edges_first):                          # boundary edges come first in a Gmsh file
            if fmt == "binary":
                f.write(struct.pack(e + "3i", 1, 1, 2) + struct.pack(e + "5i", num, 1, 1, 1, 2))
            else:
                f.write(b"%d 1 2 1 1 1 2\n" % num)
            num += 1
        if fmt == "binary":
            # elements with the same tag count share a header block; `blocks` forces extra splits
            ntag = 3 + np.diff(gp) + 1
            start = 0
            bounds = [0]
            for i in range(1, ne):
                if ntag[i] != ntag[i - 1] or (blocks > 1 and i % max(1, ne // blocks) == 0):
                    bounds.append(i)
            bounds.append(ne)
            for a, b in zip(bounds[:-1], bounds[1:]):
                f.write(struct.pack(e + "3i", 2, b - a, int(ntag[a])))
                for i in range(a, b):
                    g = gv[gp[i]:gp[i + 1]]
                    rec = [num + i, 3, 7, 1 + g.size, int(ep[i]) + 1] + [-(int(q) + 1) for q in g] + [int(v) for v in tri[i]]
                    f.write(struct.pack(e + "%di" % len(rec), *rec))
            f.write(b"\n")
        else:
            for i in range(ne):
                g = gv[gp[i]:gp[i + 1]]
                rec = [num + i, 2, 3 + g.size + 1, 3, 7, 1 + g.size, int(ep[i]) + 1] + [-(int(q) + 1) for q in g] + [int(v) for v in tri[i]]
                f.write((" ".join(map(str, rec)) + "\n").encode())
        f.write(b"$EndElements\n")


@pytest.mark.parametrize("fmt,edges_first,swap,bamg,blocks", [
    ("ascii", 0, False, False, 1), ("ascii", 5, False, True, 1), ("binary", 0, False, False, 1),
    ("binary", 7, False, False, 4), ("binary", 3, True, True, 2)])
def test_read_partitioned_msh(tmp_path, fmt, edges_first, swap, bamg, blocks):
    m = syn.make_mesh(20, 10e3, open_east=True)
    P = 4
    ep, gp, gv = tags(m, P)
    path = tmp_path / "par4mesh.msh"
    write_msh(path, m, ep, gp, gv, fmt, edges_first, swap, bamg, blocks)
    ors = orc.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv)
    for r in range(P):
        pm = capi.PartMesh.read(path, r, P, fmt=fmt, ordering="bamg" if bamg else "gmsh")
        pm.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
        # triangle numbers restart at 1 after the edges (gmshmesh.cpp:352-366), so the oracle's ids apply
        assert_matches_oracle(pm.to_local_mesh(), ors[r], m, P)


def test_reader_errors(tmp_path):
    with pytest.raises(RuntimeError, match="file not found"):
        capi.PartMesh.read(tmp_path / "missing.msh", 0, 2)
    bad = tmp_path / "bad.msh"
    bad.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 2 2 0 0 1 2 9\n$EndElements\n")
    with pytest.raises(RuntimeError, match="outside the file"):
        capi.PartMesh.read(bad, 0, 1, fmt="ascii")
    with pytest.raises(RuntimeError, match="invalid mesh file format"):
        capi.PartMesh.read(bad, 0, 1, fmt="netcdf")
    m = syn.make_mesh(8, 1.0)
    ep = np.zeros(m.ne, np.int32)                     # rank 1 owns nothing: every node still has an owner
    gp = np.zeros(m.ne + 1, np.int32)
    pm = capi.PartMesh.build(m.x, m.y, m.tri, 1, 2, ep, gp, np.zeros(0, np.int32))
    lm = pm.to_local_mesh()
    assert lm.num_nodes == 0 and lm.num_elements == 0


def test_library_is_much_faster_than_the_map_based_replay():
    """Remesh-time cost: one rank's local mesh + tables on a 200x200 mesh, 8 partitions."""
    m = syn.make_mesh(200, 10e3)
    P = 8
    ep, gp, gv = tags(m, P)
    t0 = time.perf_counter()
    ors = orc.nodal_grid(P, m.x, m.y, m.tri, ep, gp, gv, fast=True)
    for R in ors:
        R.bamg_tables()
    t_oracle = (time.perf_counter() - t0) / P
    t0 = time.perf_counter()
    pm = capi.PartMesh.build(m.x, m.y, m.tri, 3, P, ep, gp, gv)
    t_lib = time.perf_counter() - t0
    print("per rank: map-based replay %.3f s, library %.4f s" % (t_oracle, t_lib))
    assert t_lib < t_oracle

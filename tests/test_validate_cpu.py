"""nsx_create's host-side input checks (nsx_validate_mesh, no GPU): every mesh family the GPU suite uses is accepted,
malformed inputs are rejected with a message instead of being uploaded."""
import copy

import numpy as np
import pytest

from nextsim_b200 import capi, cases, partition as pt, synthetic as syn


@pytest.mark.parametrize("name,nx,nranks,open_east", [
    ("toy", None, 1, False), ("toy", None, 2, True), ("toy", None, 3, False), ("toy", None, 4, False),
    ("10km_stable", 48, 1, True), ("10km_stable", 48, 3, True), ("10km_stable", 64, 8, True), ("10km_stable", 96, 5, True),
    ("10km", 128, 2, True), ("10km", None, 1, False)])
def test_suite_meshes_are_accepted(name, nx, nranks, open_east):
    c = cases.make_case(name, nranks=nranks, nx=nx, open_east=open_east)
    for lm in c.lms:
        capi.validate_mesh(lm)


def test_random_partitions_are_accepted():
    m = syn.make_mesh(20, 10e3, open_east=True)
    ep = np.random.default_rng(1).integers(0, 6, m.ne).astype(np.int32)
    gp, gv = pt.ghost_tags(m.tri, ep, 6)
    for r in range(6):
        pm = capi.PartMesh.build(m.x, m.y, m.tri, r, 6, ep, gp, gv)
        pm.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
        capi.validate_mesh(pm.to_local_mesh())


def broken(mutate):
    c = cases.make_case("toy", nranks=2, open_east=True)
    lm = copy.deepcopy(c.lms[0])
    mutate(lm)
    return lm


@pytest.mark.parametrize("mutate,msg", [
    (lambda lm: lm.indices.__setitem__((5, 1), 0), "not a 1-based local node id"),
    (lambda lm: lm.indices.__setitem__((7, 2), lm.num_nodes + 1), "not a 1-based local node id"),
    (lambda lm: setattr(lm, "local_ndof", lm.num_nodes + 3), "local_ndof outside"),
    (lambda lm: setattr(lm, "neumann_flags", lm.neumann_flags[::-1].copy()), "sorted and unique"),
    (lambda lm: lm.nodal_connectivity.__setitem__((3, -1), 99.0), "count column out of range"),
    (lambda lm: lm.recv_from.pop(sorted(lm.recv_from)[0]), "cover every ghost node"),
    (lambda lm: lm.send_to.__setitem__(lm.rank, np.zeros(1, np.int32)), "bad send peer"),
])
def test_malformed_inputs_are_rejected(mutate, msg):
    with pytest.raises(RuntimeError, match=msg):
        capi.validate_mesh(broken(mutate))


def test_empty_rank_is_rejected():
    m = syn.make_mesh(8, 1.0)
    pm = capi.PartMesh.build(m.x, m.y, m.tri, 1, 2, np.zeros(m.ne, np.int32), np.zeros(m.ne + 1, np.int32), np.zeros(0, np.int32))
    pm.bc_marked_nodes(m.dirichlet_flags_root, m.neumann_flags_root)
    with pytest.raises(RuntimeError, match="holds no nodes or no elements"):
        capi.validate_mesh(pm.to_local_mesh())

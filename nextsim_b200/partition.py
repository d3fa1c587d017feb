"""Host-side partition indexing for the multi-GPU path (vectorised numpy).

Reproduces, from a set of per-element partition tags, exactly the local numbering the reference
derives in ``GmshMesh::nodalGrid()`` (core/src/gmshmesh.cpp:856-1498) and the exchange lists of
``FiniteElement::initUpdateGhosts()`` (model/finiteelement.cpp:14003-14088):

* rank r loads the elements with partition == r or r in the ghost list (entities.hpp:105-134);
* provisional owned nodes = nodes of owned elements (gmshmesh.cpp:907-941); duplicates between ranks go
  to the LOWEST rank, the others hold them as ghosts (1057-1094);
* local numbering = owned nodes ascending file id, then ghost nodes ascending file id (1165-1169);
* global renumbering contiguous per rank, u block then v block (1178-1220);
* kept elements: partition >= r and all three nodes local, owned first then ghosts, file order
  otherwise (1271-1312, 1384-1417);
* ghost lists grouped by owner rank in ascending reordered id, i.e. ascending file id within an owner.

Also builds the two bamg connectivity tables the solver consumes, in bamg's chain order
(contrib/bamg/src/Mesh.cpp:526-537, 583-629, 798-865): node->element lists in DESCENDING element id,
node->node lists in descending edge id (edges numbered by first appearance).

The tags themselves come from :func:`partition_elements` (recursive coordinate bisection; the
reference calls Gmsh/METIS, which is not available and whose output no reference test pins).
"""
from dataclasses import dataclass, field

import numpy as np


# ----------------------------------------------------------------------------------------------
# element -> partition tags in the msh-2.2 layout: partition p(e), ghost list G(e)
# ----------------------------------------------------------------------------------------------
def partition_elements(x, y, tri1, nparts):
    """Recursive coordinate bisection of element centroids into ``nparts`` (any count >= 1)."""
    t = tri1 - 1
    cx = x[t].mean(1)
    cy = y[t].mean(1)
    part = np.zeros(tri1.shape[0], np.int32)

    def split(ids, lo, n):
        if n == 1:
            part[ids] = lo
            return
        nl = n // 2
        spanx = np.ptp(cx[ids])
        spany = np.ptp(cy[ids])
        key = cx[ids] if spanx >= spany else cy[ids]
        k = int(round(ids.size * nl / n))
        order = np.argsort(key, kind="stable")
        split(ids[order[:k]], lo, nl)
        split(ids[order[k:]], lo + nl, n - nl)

    split(np.arange(tri1.shape[0]), 0, nparts)
    return part


def ghost_tags(tri1, elem_part, nparts):
    """G(e) = the other partitions owning an element that shares a node with e (CSR: ptr, val)."""
    ne = tri1.shape[0]
    t = (tri1 - 1).astype(np.int64)
    nn = int(t.max()) + 1
    # (node, partition) incidence, unique
    key = np.unique(t.ravel() * nparts + np.repeat(elem_part.astype(np.int64), 3))
    node = key // nparts
    prt = key % nparts
    nptr = np.zeros(nn + 1, np.int64)
    np.add.at(nptr, node + 1, 1)
    nptr = np.cumsum(nptr)
    # for every element: union over its 3 nodes of the node's partitions, minus own
    cnt = (nptr[t + 1] - nptr[t])                       # [ne,3]
    tot = cnt.sum(1)
    eid = np.repeat(np.arange(ne), tot)
    # flat gather of candidate partitions
    starts = nptr[t].ravel()
    lens = cnt.ravel()
    offs = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(lens.sum())
    cand = prt[offs]
    k2 = np.unique(eid * nparts + cand)
    e2 = k2 // nparts
    p2 = (k2 % nparts).astype(np.int32)
    keep = p2 != elem_part[e2]
    e2, p2 = e2[keep], p2[keep]
    ptr = np.zeros(ne + 1, np.int32)
    np.add.at(ptr, e2 + 1, 1)
    ptr = np.cumsum(ptr).astype(np.int32)
    return ptr, p2


# ----------------------------------------------------------------------------------------------
@dataclass
class LocalMesh:
    """What one rank's FiniteElement holds after distributedMeshProcessing (FE.cpp:50-143)."""
    rank: int
    nranks: int
    num_nodes: int                # M_num_nodes (owned + ghost)
    local_ndof: int               # M_local_ndof (owned)
    num_elements: int             # M_num_elements (owned + ghost)
    local_nelements: int          # M_local_nelements
    x: np.ndarray
    y: np.ndarray
    indices: np.ndarray           # [ne,3] int32 1-based local node ids (M_elements[].indices)
    ghostNodes: np.ndarray        # [ne,3] uint8
    node_gid: np.ndarray          # local_dof_with_ghost_init: 1-based file node ids
    node_rid: np.ndarray          # reordered global ids of (u) dofs
    elem_gid: np.ndarray          # triangles_id_with_ghost: 1-based file element numbers
    elem_part: np.ndarray
    local_ghost: np.ndarray       # sorted reordered ids of ghost nodes
    recv_from: dict = field(default_factory=dict)   # owner rank -> local ids (M_local_ghosts_local_index)
    send_to: dict = field(default_factory=dict)     # holder rank -> local ids (M_extract_local_index)
    mask_dirichlet: np.ndarray = None
    neumann_flags: np.ndarray = None
    dirichlet_flags: np.ndarray = None
    lat: np.ndarray = None
    nodal_element_connectivity: np.ndarray = None   # [nn, width] float64, NaN padded, 1-based
    nodal_connectivity: np.ndarray = None           # [nn, width+1] float64, last col = count

    @property
    def neumann_nodes(self):
        f = self.neumann_flags
        out = np.empty(2 * f.size, np.int32)
        out[0::2] = f
        out[1::2] = f + self.num_nodes
        return out


def nodal_grid(nranks, x, y, tri1, elem_part=None, ghost_ptr=None, ghost_val=None):
    """Per-rank local meshes + halo lists from partition tags (see module docstring)."""
    nn = x.size
    ne = tri1.shape[0]
    tri1 = np.ascontiguousarray(tri1, np.int32)
    if nranks == 1:
        lm = LocalMesh(rank=0, nranks=1, num_nodes=nn, local_ndof=nn, num_elements=ne, local_nelements=ne,
                       x=x.copy(), y=y.copy(), indices=tri1.copy(), ghostNodes=np.zeros((ne, 3), np.uint8),
                       node_gid=np.arange(1, nn + 1, dtype=np.int32), node_rid=np.arange(1, nn + 1, dtype=np.int32),
                       elem_gid=np.arange(1, ne + 1, dtype=np.int32), elem_part=np.zeros(ne, np.int32),
                       local_ghost=np.zeros(0, np.int32))
        return [lm]

    elem_part = np.asarray(elem_part, np.int32)
    ghost_ptr = np.asarray(ghost_ptr, np.int64)
    ghost_val = np.asarray(ghost_val, np.int32) % nranks      # entities.hpp:107-108
    nghost = np.diff(ghost_ptr)
    ghost_eid = np.repeat(np.arange(ne), nghost)
    t0 = tri1.astype(np.int64) - 1

    prov = []         # provisional owned (0-based file ids, sorted unique)
    cand_ghost = []   # all_local_nodes - prov
    loaded = []       # (element ids loaded, in file order)
    for r in range(nranks):
        own = elem_part == r
        gh = np.zeros(ne, bool)
        gh[ghost_eid[ghost_val == r]] = True
        gh &= ~own                                   # partition match wins (entities.hpp:114)
        in_gn = np.zeros(nn, bool)
        in_gn[t0[gh].ravel()] = True                 # ghosts_nodes_f
        oe = np.nonzero(own)[0]
        on = t0[oe]
        push = (~in_gn[on]) | (nghost[oe] > 0)[:, None]
        is_prov = np.zeros(nn, bool)
        is_prov[on[push]] = True
        le = np.nonzero(own | gh)[0]
        cand_e = le[(elem_part[le] >= r)]
        cand_e = cand_e[is_prov[t0[cand_e]].any(1)]
        is_local = np.zeros(nn, bool)
        is_local[t0[cand_e].ravel()] = True
        prov.append(np.nonzero(is_prov)[0])
        cand_ghost.append(np.nonzero(is_local & ~is_prov)[0])
        loaded.append(le)

    total = sum(p.size for p in prov)
    owner = np.full(nn, -1, np.int32)
    if nn < total:                                   # gmshmesh.cpp:1057-1094
        for r in range(nranks - 1, -1, -1):
            owner[prov[r]] = r                       # lowest rank wins
        owned = [np.nonzero(owner == r)[0] for r in range(nranks)]
        ghosts = [np.union1d(cand_ghost[r], np.setdiff1d(prov[r], owned[r])) for r in range(nranks)]
    else:
        owned = prov
        ghosts = cand_ghost
        for r in range(nranks):
            owner[owned[r]] = r
    if (owner < 0).any():
        raise ValueError("partition tags leave %d nodes without an owner" % int((owner < 0).sum()))

    # reorder: file id -> contiguous per-rank id (u block), 1-based
    sizes = np.array([o.size for o in owned], np.int64)
    base = 2 * np.concatenate([[0], np.cumsum(sizes)[:-1]])
    rid = np.zeros(nn, np.int64)
    for r in range(nranks):
        rid[owned[r]] = base[r] + 1 + np.arange(sizes[r])

    out = []
    for r in range(nranks):
        loc = np.concatenate([owned[r], ghosts[r]])              # local numbering (0-based file ids)
        g2l = np.full(nn, -1, np.int64)
        g2l[loc] = np.arange(loc.size)
        le = loaded[r]
        le = le[elem_part[le] >= r]
        le = le[(g2l[t0[le]] >= 0).all(1)]
        le = np.concatenate([le[elem_part[le] == r], le[elem_part[le] != r]])   # stable: owned first
        n_own_e = int((elem_part[le] == r).sum())
        ind = (g2l[t0[le]] + 1).astype(np.int32)
        gN = (ind > owned[r].size).astype(np.uint8)
        gl = np.sort(rid[ghosts[r]])
        lm = LocalMesh(rank=r, nranks=nranks, num_nodes=loc.size, local_ndof=owned[r].size,
                       num_elements=le.size, local_nelements=n_own_e,
                       x=x[loc].copy(), y=y[loc].copy(), indices=ind, ghostNodes=gN,
                       node_gid=(loc + 1).astype(np.int32), node_rid=rid[loc].astype(np.int32),
                       elem_gid=(le + 1).astype(np.int32), elem_part=elem_part[le].copy(),
                       local_ghost=gl.astype(np.int32))
        # M_local_ghosts_local_index[owner]: ghosts by ascending reordered id, grouped by owner
        g = ghosts[r]
        order = np.argsort(rid[g], kind="stable")
        g = g[order]
        go = owner[g]
        for p in np.unique(go):
            lm.recv_from[int(p)] = g2l[g[go == p]].astype(np.int32)
        out.append((lm, g, go))
    # M_extract_local_index[holder] on the owner: same nodes, owner's local ids
    for r in range(nranks):
        lm, g, go = out[r]
        for p in np.unique(go):
            sel = g[go == p]
            owner_lm = out[int(p)][0]
            # owned nodes are the first local_ndof entries in ascending file id
            own_ids = owned[int(p)]
            owner_lm.send_to[r] = np.searchsorted(own_ids, sel).astype(np.int32)
    return [o[0] for o in out]


def bc_marked_nodes(lm, dirichlet_flags_root, neumann_flags_root):
    """FiniteElement::bcMarkedNodes (FE.cpp:150-271): Dirichlet mask on OWNED nodes only; Neumann flags
    include ghosts; both as sorted unique 0-based local ids."""
    nn_glob = int(max(lm.node_gid.max(), dirichlet_flags_root.max() if dirichlet_flags_root.size else 0,
                      neumann_flags_root.max() if neumann_flags_root.size else 0)) + 1
    g2l = np.full(nn_glob, -1, np.int64)
    g2l[lm.node_gid] = np.arange(lm.num_nodes)
    d = g2l[dirichlet_flags_root]
    d = np.unique(d[(d >= 0) & (d < lm.local_ndof)])
    n = g2l[neumann_flags_root]
    n = np.unique(n[n >= 0])
    lm.dirichlet_flags = d.astype(np.int32)
    lm.neumann_flags = n.astype(np.int32)
    lm.mask_dirichlet = np.zeros(lm.num_nodes, np.uint8)
    lm.mask_dirichlet[d] = 1
    return lm


# ----------------------------------------------------------------------------------------------
# bamg connectivity tables in bamg's own ordering
# ----------------------------------------------------------------------------------------------
_VOTE = np.array([[1, 2], [2, 0], [0, 1]])      # contrib/bamg/include/macros.h:13


def bamg_tables(indices1, num_nodes):
    """(NodalElementConnectivity [nn,w] NaN padded, NodalConnectivity [nn,w2+1] with count column)."""
    t = indices1.astype(np.int64) - 1
    ne = t.shape[0]
    # node -> element, descending element id (head insertion chain, Mesh.cpp:526-537, 804-811)
    node = t.ravel()
    elem = np.repeat(np.arange(ne), 3)
    order = np.lexsort((-elem, node))
    node_s, elem_s = node[order], elem[order]
    deg = np.bincount(node, minlength=num_nodes)
    w = int(deg.max())
    start = np.concatenate([[0], np.cumsum(deg)[:-1]])
    col = np.arange(node_s.size) - start[node_s]
    nec = np.full((num_nodes, w), np.nan)
    nec[node_s, col] = elem_s + 1

    # edges numbered by first appearance over (triangle, local edge) (Mesh.cpp:583-606)
    a = t[:, _VOTE[:, 0]].ravel()
    b = t[:, _VOTE[:, 1]].ravel()
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo * num_nodes + hi
    uk, first = np.unique(key, return_index=True)
    eorder = np.argsort(first, kind="stable")            # edge id = rank of first appearance
    e_lo = (uk // num_nodes)[eorder]
    e_hi = (uk % num_nodes)[eorder]
    nedge = e_lo.size
    # node -> node, descending edge id (Mesh.cpp:830-865)
    nd = np.concatenate([e_lo, e_hi])
    other = np.concatenate([e_hi, e_lo])
    eid = np.concatenate([np.arange(nedge), np.arange(nedge)])
    order = np.lexsort((-eid, nd))
    nd_s, other_s = nd[order], other[order]
    deg2 = np.bincount(nd, minlength=num_nodes)
    w2 = int(deg2.max()) + 1
    start2 = np.concatenate([[0], np.cumsum(deg2)[:-1]])
    col2 = np.arange(nd_s.size) - start2[nd_s]
    nc = np.zeros((num_nodes, w2))
    nc[nd_s, col2] = other_s + 1
    nc[:, w2 - 1] = deg2
    return nec, nc


# ----------------------------------------------------------------------------------------------
# scatter / gather of fields between file numbering and a rank's local numbering
# ----------------------------------------------------------------------------------------------
def scatter_nodal2(lm, g, nn_glob):
    i = lm.node_gid - 1
    return np.concatenate([g[i], g[i + nn_glob]])


def scatter_nodal1(lm, g):
    return g[lm.node_gid - 1].copy()


def scatter_elem(lm, g):
    return g[..., lm.elem_gid - 1].copy()


def gather_nodal2(lms, locs, nn_glob):
    out = np.full(2 * nn_glob, np.nan)
    for lm, v in zip(lms, locs):
        i = lm.node_gid[:lm.local_ndof] - 1
        out[i] = v[:lm.local_ndof]
        out[i + nn_glob] = v[lm.num_nodes:lm.num_nodes + lm.local_ndof]
    return out


def gather_elem(lms, locs, ne_glob):
    out = np.full(ne_glob, np.nan)
    for lm, v in zip(lms, locs):
        out[lm.elem_gid[:lm.local_nelements] - 1] = v[:lm.local_nelements]
    return out

// nsx_mesh.cpp -- builds the MeshPlan (see nsx_mesh.h) from the reference-side mesh description.
//
// Internal numbering: owned nodes sorted along a Hilbert curve through their coordinates (ghost nodes after,
// also Hilbert-sorted, so "id >= local_ndof <=> ghost" still holds); elements grouped by writer tile.  The
// host never sees this numbering: nsx_upload / nsx_download permute on the device.  Summation orders that the
// reference fixes (ascending element id for the stress gradient, bamg chain order for tau_a and the OW
// smoother) are preserved by storing the incidence lists in REFERENCE order with internal ids as payload.
#include "nsx_mesh.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace nsx {

static uint64_t hilbert_xy2d(uint32_t x, uint32_t y)        // 16-bit coordinates -> 32-bit curve index
{
    uint32_t const n = 1u << 16;
    uint64_t d = 0;
    for (uint32_t s = n / 2; s > 0; s /= 2) {
        uint32_t const rx = (x & s) > 0, ry = (y & s) > 0;
        d += (uint64_t)s * s * ((3 * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = n - 1 - x; y = n - 1 - y; }
            std::swap(x, y);
        }
    }
    return d;
}

void build_mesh_plan(const NsxMesh* M, MeshPlan& P, int target_tile_nodes, int sm_count,
                     bool resident_order, const uint8_t* export_mask, int max_tile_nodes)
{
    int const nn = M->num_nodes, ne = M->num_elements, ndof = M->local_ndof;
    if (nn <= 0 || ne <= 0 || ndof <= 0 || ndof > nn || M->local_nelements > ne)
        throw std::invalid_argument("nsx_create: inconsistent mesh sizes");
    if (!M->coord_x || !M->coord_y || !M->indices || !M->lat || !M->nodal_element_connectivity || !M->nodal_connectivity)
        throw std::invalid_argument("nsx_create: NULL mesh array");
    if ((long)3 * ne >= (1L << 31)) throw std::invalid_argument("nsx_create: mesh too large for 32-bit slot ids");
    P.nn = nn; P.ndof = ndof; P.ne = ne; P.ne_local = M->local_nelements;

    // ---- reference connectivity, validated --------------------------------------------------------------
    std::vector<int> r0(ne), r1(ne), r2(ne), rdeg(nn, 0);
    for (int e = 0; e < ne; ++e) {
        int const a = M->indices[3 * (size_t)e] - 1, b = M->indices[3 * (size_t)e + 1] - 1, c = M->indices[3 * (size_t)e + 2] - 1;
        if (a < 0 || b < 0 || c < 0 || a >= nn || b >= nn || c >= nn)
            throw std::invalid_argument("nsx_create: element index out of range");
        if (M->ghost_nodes) {       // GMSHElement::ghostNodes == "local id >= local_ndof" (gmshmesh.cpp:1298-1301)
            const unsigned char* g = M->ghost_nodes + 3 * (size_t)e;
            if ((g[0] != 0) != (a >= ndof) || (g[1] != 0) != (b >= ndof) || (g[2] != 0) != (c >= ndof))
                throw std::invalid_argument("nsx_create: ghost_nodes disagrees with the owned-first node numbering");
        }
        r0[e] = a; r1[e] = b; r2[e] = c;
        rdeg[a]++; rdeg[b]++; rdeg[c]++;
    }
    int w = 0;
    for (int n = 0; n < nn; ++n) {
        if (rdeg[n] == 0) throw std::invalid_argument("nsx_create: orphan node");
        w = std::max(w, rdeg[n]);
    }
    P.ell_w = w;

    // ---- node permutation: Hilbert order, owned first ---------------------------------------------------
    // bounding box of the OWNED nodes, each axis stretched to the full curve square: a domain edge that pokes a
    // little beyond a major quadrant line would otherwise be cut off as a one-node-wide sliver tile
    double xmin = M->coord_x[0], xmax = xmin, ymin = M->coord_y[0], ymax = ymin;
    for (int n = 1; n < ndof; ++n) {
        xmin = std::min(xmin, M->coord_x[n]); xmax = std::max(xmax, M->coord_x[n]);
        ymin = std::min(ymin, M->coord_y[n]); ymax = std::max(ymax, M->coord_y[n]);
    }
    double const spanx = std::max(xmax - xmin, 1e-300), spany = std::max(ymax - ymin, 1e-300);
    std::vector<uint64_t> key(nn);
    for (int n = 0; n < nn; ++n) {
        double const fx = std::min(1.0, std::max(0.0, (M->coord_x[n] - xmin) / spanx));
        double const fy = std::min(1.0, std::max(0.0, (M->coord_y[n] - ymin) / spany));
        key[n] = hilbert_xy2d((uint32_t)(fx * 65535.0), (uint32_t)(fy * 65535.0));
    }
    P.node_inv.resize(nn);
    std::iota(P.node_inv.begin(), P.node_inv.end(), 0);
    auto by_key = [&](int a, int b) { return key[a] != key[b] ? key[a] < key[b] : a < b; };
    std::sort(P.node_inv.begin(), P.node_inv.begin() + ndof, by_key);
    std::sort(P.node_inv.begin() + ndof, P.node_inv.end(), by_key);
    // ---- tiles over the owned nodes ------------------------------------------------------------------------
    int T = std::max(32, target_tile_nodes);
    int ntiles = (ndof + T - 1) / T;
    // `sm_count` here is the number of CTAs of one full wave (SMs x resident CTAs per SM): small meshes get a
    // whole number of waves so that no SM idles during a partial last wave
    if (sm_count > 0 && ntiles < 8 * sm_count) {
        int const nwaves = std::max(1, (ntiles + sm_count / 2) / sm_count);
        ntiles = nwaves * sm_count;
    }
    ntiles = std::max(1, std::min(ntiles, ndof));
    T = (ndof + ntiles - 1) / ntiles;
    ntiles = (ndof + T - 1) / T;
    // tile t owns the nodes [cut[t], cut[t+1]) of the curve order.  Equal node counts by default; the resident solver
    // (one tile per SM for a whole model step, the slowest tile sets the pace) balances the SLOT counts instead: a few
    // rounds of "measure the slots of every tile, move the cuts to equalise the cost along the curve".
    std::vector<int> cut(ntiles + 1);
    for (int t = 0; t <= ntiles; ++t) cut[t] = std::min(ndof, t * T);
    if (resident_order && ntiles > 1) {
        std::vector<std::vector<int>> inc_of(ndof);            // reference node -> incident elements
        for (int e = 0; e < ne; ++e) {
            int const v[3] = {r0[e], r1[e], r2[e]};
            for (int i = 0; i < 3; ++i) if (v[i] < ndof) inc_of[v[i]].push_back(e);
        }
        std::vector<int> stamp(ne, -1);
        int const max_nodes = std::max(T, max_tile_nodes);
        // slot_cap: two slots per thread (a third round of the slot loop costs every tile that needs it a full extra
        // slot latency): tiles above it are made a little more expensive each extra round until they fit, if they can
        int const slot_cap = 2 * max_tile_nodes;
        std::vector<double> penalty(ntiles, 1.0);
        for (int round = 0; round < 24; ++round) {
            std::vector<double> cost(ndof);                      // per curve position: slots of its tile / nodes of its tile
            int worst = 0;
            for (int t = 0; t < ntiles; ++t) {
                int slots = 0;
                for (int i = cut[t]; i < cut[t + 1]; ++i)
                    for (int e : inc_of[P.node_inv[i]]) if (stamp[e] != round * ntiles + t) { stamp[e] = round * ntiles + t; ++slots; }
                worst = std::max(worst, slots);
                if (round >= 6 && max_tile_nodes > 0 && slots > slot_cap) penalty[t] *= 1.02;
                double const c = penalty[t] * (double)slots / std::max(1, cut[t + 1] - cut[t]);
                for (int i = cut[t]; i < cut[t + 1]; ++i) cost[i] = c;
            }
            if (round >= 6 && (max_tile_nodes <= 0 || worst <= slot_cap)) break;
            double total = 0.;
            for (int i = 0; i < ndof; ++i) total += cost[i];
            std::vector<int> nc(ntiles + 1, 0);
            nc[ntiles] = ndof;
            double acc = 0.;
            int t = 1;
            for (int i = 0; i < ndof && t < ntiles; ++i) {
                acc += cost[i];
                while (t < ntiles && acc >= total * t / ntiles) nc[t++] = i + 1;
            }
            for (; t < ntiles; ++t) nc[t] = ndof;
            bool ok = true;                                      // every tile non-empty and within the thread budget
            for (int q = 0; q < ntiles; ++q) if (nc[q + 1] <= nc[q] || nc[q + 1] - nc[q] > max_nodes) ok = false;
            if (!ok) break;
            cut = nc;
        }
    }
    T = 0;
    for (int t = 0; t < ntiles; ++t) T = std::max(T, cut[t + 1] - cut[t]);
    P.ntiles = ntiles; P.tile_nodes = T;                         // tile_nodes = the LARGEST tile
    P.tile_of.assign(ndof, 0);
    for (int t = 0; t < ntiles; ++t) for (int i = cut[t]; i < cut[t + 1]; ++i) P.tile_of[i] = t;
    auto tile_of_node = [&](int internal) { return P.tile_of[internal]; };

    // resident solver: inside every tile the EXPORT nodes come first -- owned nodes that another tile reads (they share
    // an element with a node of another tile) or that are sent to another rank.  Tile membership does not change.
    std::vector<uint8_t> is_x;
    if (resident_order) {
        std::vector<int> pos(nn);
        for (int i = 0; i < nn; ++i) pos[P.node_inv[i]] = i;
        is_x.assign(nn, 0);                                   // by reference id
        for (int e = 0; e < ne; ++e) {
            int const v[3] = {r0[e], r1[e], r2[e]};
            for (int i = 0; i < 3; ++i) {
                if (v[i] >= ndof) continue;
                int const t = P.tile_of[pos[v[i]]];
                for (int j = 0; j < 3; ++j)
                    if (j != i && v[j] < ndof && P.tile_of[pos[v[j]]] != t) is_x[v[i]] = 1;
            }
        }
        if (export_mask)
            for (int r = 0; r < ndof; ++r) if (export_mask[r]) is_x[r] = 1;
        for (int t = 0; t < ntiles; ++t) {
            auto b = P.node_inv.begin() + cut[t], e = P.node_inv.begin() + cut[t + 1];
            std::stable_partition(b, e, [&](int r) { return is_x[r] != 0; });
        }
    }
    P.node_perm.resize(nn);
    for (int i = 0; i < nn; ++i) P.node_perm[P.node_inv[i]] = i;

    P.x.resize(nn); P.y.resize(nn); P.lat.resize(nn); P.nflags.assign(nn, 0);
    for (int i = 0; i < nn; ++i) {
        int const r = P.node_inv[i];
        P.x[i] = M->coord_x[r]; P.y[i] = M->coord_y[r]; P.lat[i] = M->lat[r];
        if (M->mask_dirichlet && M->mask_dirichlet[r]) P.nflags[i] |= 1;      // NF_DIRICHLET
        if (r >= ndof) P.nflags[i] |= 4;                                        // NF_GHOST
        if (std::signbit(M->lat[r])) P.nflags[i] |= 8;                          // NF_LATNEG
    }
    for (int k = 0; k < M->n_neumann_flags; ++k) {
        int const r = M->neumann_flags[k];
        if (r < 0 || r >= nn) throw std::invalid_argument("nsx_create: neumann flag out of range");
        P.nflags[P.node_perm[r]] |= 2;                                          // NF_NEUMANN
    }

    // ---- element permutation: by writer tile, then along the curve ------------------------------------------
    std::vector<int> writer(ne, -1);
    std::vector<uint64_t> ekey(ne);
    for (int e = 0; e < ne; ++e) {
        int const v[3] = {P.node_perm[r0[e]], P.node_perm[r1[e]], P.node_perm[r2[e]]};
        int lo = nn;
        for (int i = 0; i < 3; ++i) if (v[i] < ndof) lo = std::min(lo, v[i]);
        if (lo < nn) writer[e] = tile_of_node(lo);
    }
    // ghost elements without an owned node: written by a tile that already reads one of their (ghost) nodes, so that
    // no interior tile starts depending on another rank
    {
        std::vector<int> tile_of_ghost(nn, -1);
        for (int e = 0; e < ne; ++e) {
            if (writer[e] < 0) continue;
            int const v[3] = {r0[e], r1[e], r2[e]};
            for (int i = 0; i < 3; ++i) if (v[i] >= ndof && tile_of_ghost[v[i]] < 0) tile_of_ghost[v[i]] = writer[e];
        }
        for (int e = 0; e < ne; ++e) {
            if (writer[e] >= 0) continue;
            int const v[3] = {r0[e], r1[e], r2[e]};
            for (int i = 0; i < 3 && writer[e] < 0; ++i) writer[e] = tile_of_ghost[v[i]];
            if (writer[e] < 0) writer[e] = e % ntiles;
        }
    }
    std::vector<uint8_t> late(ne, 0);              // resident order: all three nodes interior to the writer tile
    for (int e = 0; e < ne; ++e) {
        int const v[3] = {P.node_perm[r0[e]], P.node_perm[r1[e]], P.node_perm[r2[e]]};
        if (resident_order) {
            int const rv[3] = {r0[e], r1[e], r2[e]};
            bool interior = true;
            for (int i = 0; i < 3; ++i)
                interior = interior && rv[i] < ndof && tile_of_node(v[i]) == writer[e] && !is_x[rv[i]];
            late[e] = interior ? 1 : 0;
        }
        ekey[e] = ((uint64_t)writer[e] << 32) | ((uint64_t)late[e] << 31) | (uint32_t)std::min(std::min(v[0], v[1]), v[2]);
    }
    // Resident solver: CONFLICT-AWARE SLOT PLACEMENT.  In the nodal solve the 16 threads of a half-warp (16 consecutive
    // owned nodes) read, for incidence column c, the stress / constants of 16 different slots with 64-bit shared-memory
    // loads: conflict-free exactly when the 16 slot indices differ modulo 16 (the planes are a multiple of 16 apart).
    // Slots may be placed freely inside their group (early own | late own | halo), so each slot greedily takes the
    // residue that collides least in the (half-warp, column) cells it belongs to, within the capacity of its group.
    std::vector<std::vector<int>> halo_order;
    if (resident_order) {
        halo_order.assign(ntiles, {});
        std::vector<std::vector<int>> inc_ref(ndof);             // reference node -> incident elements, ascending
        for (int e = 0; e < ne; ++e) {
            int const v[3] = {r0[e], r1[e], r2[e]};
            for (int i = 0; i < 3; ++i) if (v[i] < ndof) inc_ref[v[i]].push_back(e);
        }
        std::vector<std::vector<int>> own_of(ntiles);
        for (int e = 0; e < ne; ++e) own_of[writer[e]].push_back(e);
        std::vector<int> grp(ne, -1), seen(ne, -1), res(ne, 0), ncell(ne, 0);
        std::vector<int> cell_of((size_t)ne * 3, 0);
        long wf = 0, ideal = 0;
        for (int t = 0; t < ntiles; ++t) {
            int const no = cut[t + 1] - cut[t];
            std::vector<int> members[3];
            for (int e : own_of[t]) { grp[e] = late[e] ? 1 : 0; members[grp[e]].push_back(e); seen[e] = t; ncell[e] = 0; }
            int wmax = 0;
            for (int pj = 0; pj < no; ++pj) {
                auto const& L = inc_ref[P.node_inv[cut[t] + pj]];
                wmax = std::max(wmax, (int)L.size());
                for (int e : L) if (seen[e] != t) { seen[e] = t; grp[e] = 2; ncell[e] = 0; members[2].push_back(e); }
            }
            int const nhw = (no + 15) / 16, ncells = nhw * std::max(1, wmax);
            std::vector<uint8_t> cnt((size_t)ncells * 16, 0);
            std::vector<int> order;                               // slots in the order the node loop meets them
            order.reserve(members[0].size() + members[1].size() + members[2].size());
            std::vector<int> queued(0);
            for (int pj = 0; pj < no; ++pj) {
                auto const& L = inc_ref[P.node_inv[cut[t] + pj]];
                for (int c = 0; c < (int)L.size(); ++c) {
                    int const e = L[c];
                    if (ncell[e] == 0) order.push_back(e);
                    if (ncell[e] < 3) cell_of[(size_t)e * 3 + ncell[e]++] = (pj / 16) * wmax + c;
                }
            }
            for (int g = 0; g < 3; ++g) for (int e : members[g]) if (ncell[e] == 0) order.push_back(e);   // no owned node of this tile
            int const base[4] = {0, (int)members[0].size(), (int)(members[0].size() + members[1].size()),
                                 (int)(members[0].size() + members[1].size() + members[2].size())};
            int cap[3][16];
            for (int g = 0; g < 3; ++g) for (int r = 0; r < 16; ++r) cap[g][r] = 0;
            for (int g = 0; g < 3; ++g) for (int q = base[g]; q < base[g + 1]; ++q) cap[g][q & 15]++;
            for (int e : order) {
                int const g = grp[e];
                int best = -1, best_cost = 1 << 30, best_cap = -1;
                for (int r = 0; r < 16; ++r) {
                    if (cap[g][r] <= 0) continue;
                    int cost = 0;
                    for (int q = 0; q < ncell[e]; ++q) cost += cnt[(size_t)cell_of[(size_t)e * 3 + q] * 16 + r];
                    if (cost < best_cost || (cost == best_cost && cap[g][r] > best_cap)) { best = r; best_cost = cost; best_cap = cap[g][r]; }
                }
                res[e] = best;
                cap[g][best]--;
                for (int q = 0; q < ncell[e]; ++q) cnt[(size_t)cell_of[(size_t)e * 3 + q] * 16 + best]++;
            }
            // positions: inside each group the free positions of residue r are handed out in placement order
            int nextpos[3][16];
            for (int g = 0; g < 3; ++g)
                for (int r = 0; r < 16; ++r) {
                    int q = base[g];
                    while ((q & 15) != r) ++q;
                    nextpos[g][r] = q;
                }
            std::vector<int> halo_pos(members[2].size());
            std::vector<std::pair<int, int>> hp;
            for (int e : order) {
                int const g = grp[e], q = nextpos[g][res[e]];
                nextpos[g][res[e]] += 16;
                if (q >= base[g + 1]) throw std::logic_error("nsx mesh plan: slot placement ran out of positions");
                if (g < 2) ekey[e] = ((uint64_t)writer[e] << 32) | (uint32_t)q;
                else hp.emplace_back(q, e);
            }
            std::sort(hp.begin(), hp.end());
            for (auto const& pe : hp) halo_order[t].push_back(pe.second);
            for (int cl = 0; cl < ncells; ++cl) {
                int mx = 0;
                for (int r = 0; r < 16; ++r) mx = std::max(mx, (int)cnt[(size_t)cl * 16 + r]);
                wf += mx; ideal += mx > 0;
            }
        }
        P.p2_wavefronts = wf; P.p2_cells = ideal;
    }
    P.elem_inv.resize(ne);
    std::iota(P.elem_inv.begin(), P.elem_inv.end(), 0);
    std::sort(P.elem_inv.begin(), P.elem_inv.end(), [&](int a, int b) { return ekey[a] != ekey[b] ? ekey[a] < ekey[b] : a < b; });
    P.elem_perm.resize(ne);
    for (int i = 0; i < ne; ++i) P.elem_perm[P.elem_inv[i]] = i;
    for (int i = 0; i < 3; ++i) P.en[i].resize(ne);
    std::vector<int> iwriter(ne);
    for (int i = 0; i < ne; ++i) {
        int const r = P.elem_inv[i];
        P.en[0][i] = P.node_perm[r0[r]]; P.en[1][i] = P.node_perm[r1[r]]; P.en[2][i] = P.node_perm[r2[r]];
        iwriter[i] = writer[r];
    }

    // ---- node -> element ELL in ascending reference element order ---------------------------------------------
    P.n2e.assign((size_t)w * nn, -1);
    P.n2e_deg.assign(nn, 0);
    for (int r = 0; r < ne; ++r) {
        int const v[3] = {r0[r], r1[r], r2[r]};
        int const ie = P.elem_perm[r];
        for (int i = 0; i < 3; ++i) {
            int const n = P.node_perm[v[i]];
            P.n2e[(size_t)P.n2e_deg[n] * nn + n] = i * ne + ie;
            P.n2e_deg[n]++;
        }
    }

    // ---- bamg tables -> int ELL in the given order (quirks Q6, Q7) ---------------------------------------------
    int const nw = M->nec_width;
    P.nec_w = nw;
    P.nec.assign((size_t)std::max(nw, 1) * nn, -1);
    for (int r = 0; r < nn; ++r)
        for (int j = 0; j < nw; ++j) {
            double const raw = M->nodal_element_connectivity[(size_t)nw * r + j];
            int e = -1;
            if (!std::isnan(raw)) e = (int)raw - 1;
            if (e >= ne) throw std::invalid_argument("nsx_create: NodalElementConnectivity entry out of range");
            P.nec[(size_t)j * nn + P.node_perm[r]] = e < 0 ? -1 : P.elem_perm[e];
        }
    int const cw = M->nc_width;
    if (cw < 1) throw std::invalid_argument("nsx_create: bad NodalConnectivity width");
    P.nc_w = cw - 1;
    P.n2n.assign((size_t)std::max(cw - 1, 1) * nn, 0);
    P.n2n_deg.assign(nn, 0);
    for (int r = 0; r < nn; ++r) {
        int const cnt = (int)M->nodal_connectivity[(size_t)cw * (r + 1) - 1];
        if (cnt < 0 || cnt > cw - 1) throw std::invalid_argument("nsx_create: bad NodalConnectivity count");
        int const n = P.node_perm[r];
        P.n2n_deg[n] = cnt;
        for (int j = 0; j < cnt; ++j) {
            int const q = (int)M->nodal_connectivity[(size_t)cw * r + j] - 1;
            if (q < 0 || q >= nn) throw std::invalid_argument("nsx_create: NodalConnectivity entry out of range");
            P.n2n[(size_t)j * nn + n] = P.node_perm[q];
        }
    }

    // ---- per-tile structures ------------------------------------------------------------------------------------
    std::vector<int> own_cnt(ntiles, 0), own_begin(ntiles + 1, 0);
    for (int i = 0; i < ne; ++i) own_cnt[iwriter[i]]++;
    for (int t = 0; t < ntiles; ++t) own_begin[t + 1] = own_begin[t] + own_cnt[t];
    int const nghost = nn - ndof;
    int const G = (nghost + ntiles - 1) / ntiles;

    P.tiles.assign(ntiles, TileDesc{});
    P.halo_nodes.clear(); P.halo_elems.clear(); P.slot_elem.clear(); P.slot_conn.clear(); P.inc.clear();
    P.res_tiles.clear(); P.res_nbr.clear(); P.res_n2n.clear(); P.res_n2n_deg.clear(); P.halo_move.clear();
    if (resident_order) { P.res_tiles.assign(ntiles, ResTile{}); P.res_n2n_deg.assign(ndof, 0); }
    std::vector<uint8_t> ghost_taken(nn, 0);
    std::vector<int> nbr_stamp(ntiles, -1);
    std::vector<int> stamp_e(ne, -1), slot_of(ne, 0), stamp_n(nn, -1), lidx(nn, 0);
    P.max_local_nodes = 0; P.max_slots = 0; P.max_own_slots = 0; P.max_halo_slots = 0; P.max_halo_nodes = 0; P.max_inc = 0;
    for (int t = 0; t < ntiles; ++t) {
        TileDesc& td = P.tiles[t];
        td.node_begin = cut[t];
        td.n_own = cut[t + 1] - cut[t];
        td.elem_begin = own_begin[t];
        td.n_own_slots = own_cnt[t];
        td.slot_begin = (int)P.slot_elem.size();
        td.halo_off = (int)P.halo_nodes.size();
        td.halo_elem_off = (int)P.halo_elems.size();
        td.ghost_begin = ndof + std::min(nghost, t * G);
        td.n_ghost = std::min(nghost, (t + 1) * G) - std::min(nghost, t * G);
        // own slots
        for (int k = 0; k < td.n_own_slots; ++k) {
            int const ie = td.elem_begin + k;
            stamp_e[ie] = t; slot_of[ie] = k;
            P.slot_elem.push_back(ie);
        }
        // halo slots: elements touching my owned nodes but written by another tile (reference order per node)
        int nh = 0, dmax = 0;
        if (resident_order)                                  // placed by the conflict-aware pass above
            for (int re : halo_order[t]) {
                int const ie = P.elem_perm[re];
                stamp_e[ie] = t; slot_of[ie] = td.n_own_slots + nh; ++nh;
                P.halo_elems.push_back(ie);
                P.slot_elem.push_back(ie);
            }
        for (int j = 0; j < td.n_own; ++j) {
            int const n = td.node_begin + j;
            dmax = std::max(dmax, P.n2e_deg[n]);
            for (int c = 0; c < P.n2e_deg[n]; ++c) {
                int const ie = P.n2e[(size_t)c * nn + n] % ne;
                if (stamp_e[ie] != t) {
                    if (resident_order) throw std::logic_error("nsx mesh plan: halo slot missing from the placement pass");
                    stamp_e[ie] = t; slot_of[ie] = td.n_own_slots + nh; ++nh;
                    P.halo_elems.push_back(ie);
                    P.slot_elem.push_back(ie);
                }
            }
        }
        td.n_halo_slots = nh;
        int const nslots = td.n_own_slots + nh;
        // local nodes: owned range first, then every other node of the tile's elements
        for (int j = 0; j < td.n_own; ++j) { stamp_n[td.node_begin + j] = t; lidx[td.node_begin + j] = j; }
        int nhn = 0;
        for (int k = 0; k < nslots; ++k) {
            int const ie = P.slot_elem[td.slot_begin + k];
            unsigned long long packed = 0;
            for (int i = 0; i < 3; ++i) {
                int const n = P.en[i][ie];
                if (stamp_n[n] != t) {
                    stamp_n[n] = t; lidx[n] = td.n_own + HALO_GAP + nhn; ++nhn;
                    P.halo_nodes.push_back(n);
                    if (n >= ndof) td.boundary = 1;          // reads a ghost slot
                }
                packed |= (unsigned long long)(lidx[n] & 0xFFFF) << (16 * i);
            }
            P.slot_conn.push_back(packed);
        }
        td.n_halo = nhn;
        if (td.n_own + HALO_GAP + nhn > 65535 || 3 * nslots > 65534)
            throw std::invalid_argument("nsx_create: tile too large for 16-bit local ids");
        // incidence table of the owned nodes, reference-ascending element order, column-major
        td.inc_off = (int)P.inc.size();
        td.inc_w = dmax;
        P.inc.resize(P.inc.size() + (size_t)dmax * td.n_own, (uint16_t)0xFFFF);
        for (int j = 0; j < td.n_own; ++j) {
            int const n = td.node_begin + j;
            for (int c = 0; c < P.n2e_deg[n]; ++c) {
                int const s = P.n2e[(size_t)c * nn + n];
                int const i = s / ne, ie = s - i * ne;
                P.inc[td.inc_off + (size_t)c * td.n_own + j] = (uint16_t)(slot_of[ie] * 3 + i);
            }
        }
        if (resident_order) {
            ResTile& rt = P.res_tiles[t];
            for (int j = 0; j < td.n_own; ++j) if (is_x[P.node_inv[td.node_begin + j]]) rt.n_x = j + 1;
            for (int j = 0; j < rt.n_x; ++j)
                if (!is_x[P.node_inv[td.node_begin + j]]) throw std::logic_error("nsx mesh plan: export nodes are not a prefix");
            for (int k = 0; k < td.n_own_slots; ++k) {
                bool const lt = late[P.elem_inv[td.elem_begin + k]] != 0;
                if (!lt) { if (rt.n_early_own != k) throw std::logic_error("nsx mesh plan: early slots are not a prefix"); rt.n_early_own = k + 1; }
            }
            // neighbour tiles = owners of my halo nodes (the relation is symmetric: the shared element is a slot of both)
            rt.nbr_off = (int)P.res_nbr.size();
            for (int h = 0; h < nhn; ++h) {
                int const g = P.halo_nodes[td.halo_off + h];
                uint8_t mv = 0;
                if (g >= ndof) { if (!ghost_taken[g]) { ghost_taken[g] = 1; mv = 1; } }
                else {
                    int const q = tile_of_node(g);
                    if (nbr_stamp[q] != t) { nbr_stamp[q] = t; P.res_nbr.push_back(q); }
                }
                P.halo_move.push_back(mv);
            }
            rt.n_nbr = (int)P.res_nbr.size() - rt.nbr_off;
            // node -> node table of the owned nodes in tile-local ids, bamg order (open-water smoother, FE.cpp:10597-10605)
            int wmax = 0;
            for (int j = 0; j < td.n_own; ++j) wmax = std::max(wmax, P.n2n_deg[td.node_begin + j]);
            rt.n2n_off = (int)P.res_n2n.size();
            rt.n2n_w = wmax;
            P.res_n2n.resize(P.res_n2n.size() + (size_t)wmax * td.n_own, (uint16_t)0);
            for (int j = 0; j < td.n_own; ++j) {
                int const n = td.node_begin + j;
                P.res_n2n_deg[n] = (uint8_t)P.n2n_deg[n];
                for (int c = 0; c < P.n2n_deg[n]; ++c) {
                    int const q = P.n2n[(size_t)c * nn + n];
                    if (stamp_n[q] != t) throw std::invalid_argument("nsx_create: NodalConnectivity lists a node that shares no element with its row");
                    P.res_n2n[rt.n2n_off + (size_t)c * td.n_own + j] = (uint16_t)lidx[q];
                }
            }
        }
        P.max_local_nodes = std::max(P.max_local_nodes, td.n_own + HALO_GAP + nhn);
        P.max_slots = std::max(P.max_slots, nslots);
        P.max_own_slots = std::max(P.max_own_slots, td.n_own_slots);
        P.max_halo_slots = std::max(P.max_halo_slots, nh);
        P.max_halo_nodes = std::max(P.max_halo_nodes, nhn);
        P.max_inc = std::max(P.max_inc, dmax * td.n_own);
    }
    if (resident_order) {
        // mailbox slots: export nodes in tile order, then the ghost nodes
        int off = 0;
        for (int t = 0; t < ntiles; ++t) { P.res_tiles[t].x_off = off; off += P.res_tiles[t].n_x; }
        P.n_export = off;
        P.halo_slot.assign(P.halo_nodes.size(), 0);
        for (size_t h = 0; h < P.halo_nodes.size(); ++h) {
            int const g = P.halo_nodes[h];
            if (g >= ndof) { P.halo_slot[h] = off + (g - ndof); continue; }
            int const q = tile_of_node(g), j = g - cut[q];
            if (j >= P.res_tiles[q].n_x) throw std::logic_error("nsx mesh plan: a halo node is not an export node of its tile");
            P.halo_slot[h] = P.res_tiles[q].x_off + j;
        }
    }
    // slot space is padded to an even count so that every slot plane (stride nslots) has the same 16-byte
    // phase; the pad slot is never computed (marked INT_MIN)
    if (P.slot_elem.size() & 1) { P.slot_elem.push_back(INT32_MIN); P.slot_conn.push_back(0ULL); }
    P.nslots = (int)P.slot_elem.size();

    // post-pass: incidence codes become shared-memory offsets (vertex * msp + slot); halo slots get ~e
    P.msp = (P.max_slots + 3) & ~1;             // plane stride in shared memory: room for the alignment shift
    if (resident_order) P.msp = (P.max_slots + 15) & ~15;      // planes a multiple of 16 slots apart: bank = slot mod 16 in every plane
    int const MS = P.msp;
    for (int t = 0; t < ntiles; ++t) {
        TileDesc const& td = P.tiles[t];
        for (size_t q = 0; q < (size_t)td.inc_w * td.n_own; ++q) {
            uint16_t& code = P.inc[td.inc_off + q];
            if (code == 0xFFFF) continue;
            int const slot = code / 3, i = code % 3;
            code = (uint16_t)(i * MS + slot);
        }
        for (int k = td.n_own_slots; k < td.n_own_slots + td.n_halo_slots; ++k)
            P.slot_elem[td.slot_begin + k] = ~P.slot_elem[td.slot_begin + k];
    }
    if (6 * (long)MS > 65534) throw std::invalid_argument("nsx_create: tile too large for 16-bit incidence codes");
}

}  // namespace nsx

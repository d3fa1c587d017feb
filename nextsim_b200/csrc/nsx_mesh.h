// nsx_mesh.h -- host-side mesh plan: internal (space-filling-curve) numbering and the tile decomposition the
// fused sub-cycle kernel works on.  Pure host C++ (no CUDA types); built once per nsx_create, i.e. per remesh.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/nsx.h"

namespace nsx {

// One tile = one CTA of the sub-cycle kernel: a contiguous range of owned nodes (internal numbering) plus
// every element touching them.  Elements whose writer is another tile are recomputed redundantly ("halo
// slots", ~2/sqrt(T) of the tile) exactly like the reference recomputes ghost elements on the lower rank.
struct TileDesc {
    int node_begin, n_own;          // owned nodes [node_begin, node_begin + n_own)
    int halo_off, n_halo;           // other nodes read by the tile: halo_nodes[halo_off ...]
    int slot_begin;                 // first slot of the tile in slot space
    int n_own_slots, n_halo_slots;  // own slots map to elements [elem_begin + k]; halo slots via halo_elems
    int elem_begin;
    int halo_elem_off;
    int inc_off, inc_w;             // incidence table of the owned nodes: inc[inc_off + c*n_own + j]
    int ghost_begin, n_ghost;       // ghost nodes whose (lagged) mesh move this tile performs
    int boundary;                   // 1 if one of its owned nodes is sent to another rank
    int pad_[2];
};
static_assert(sizeof(TileDesc) == 64, "TileDesc is 16 ints");

// Extra per-tile data of the state-resident solver (k_resident): one tile per SM for the whole sub-cycle loop.
// Inside a tile the owned nodes are ordered EXPORT NODES FIRST (nodes another tile reads as halo, or that are sent to
// another rank) and the own slots EARLY SLOTS FIRST (elements with a node outside the tile's interior, i.e. an export
// node or a halo node); halo slots are always early.  Per sub-cycle a tile can then finish its export nodes, publish
// them, and hide the neighbours' latency behind the interior work.
struct ResTile {
    int n_x;                        // owned nodes [0, n_x) of the tile are export nodes
    int n_early_own;                // own slots [0, n_early_own) are early
    int nbr_off, n_nbr;             // neighbour tiles (owners of my halo nodes, symmetric): res_nbr[nbr_off ...]
    int n2n_off, n2n_w;             // bamg-order node->node table of the owned nodes in tile-local ids (smoother)
    int x_off;                      // mailbox slot of the tile's first export node (prefix sum of n_x over the tiles)
    int pad_;
};
static_assert(sizeof(ResTile) == 32, "ResTile is 8 ints");

// The owned nodes' velocities are staged by a TMA copy that is rounded up to 16-byte granules, i.e. it may write up to
// two doubles past the owned range; the halo nodes (written by plain stores) therefore start HALO_GAP entries later.
constexpr int HALO_GAP = 2;

struct MeshPlan {
    int nn = 0, ndof = 0, ne = 0, ne_local = 0;
    std::vector<int> node_perm, node_inv;      // reference local id -> internal id, and back
    std::vector<int> elem_perm, elem_inv;
    std::vector<double> x, y, lat;             // internal numbering
    std::vector<uint8_t> nflags;
    std::vector<int> en[3];                    // internal element -> internal node
    // node -> (element, vertex) in ASCENDING REFERENCE element order (the order of FE.cpp:10445-10467),
    // column-major ELL, value = vertex * ne + internal element id
    int ell_w = 0;
    std::vector<int> n2e, n2e_deg;
    int nec_w = 0;                             // bamg NodalElementConnectivity order (FE.cpp:10376-10390)
    std::vector<int> nec;
    int nc_w = 0;                              // bamg NodalConnectivity order (FE.cpp:10597-10605)
    std::vector<int> n2n, n2n_deg;
    // tiles
    int ntiles = 0, tile_nodes = 0 /* nodes of the largest tile */, nslots = 0, max_local_nodes = 0, max_slots = 0;
    int max_own_slots = 0, max_halo_slots = 0, max_halo_nodes = 0, max_inc = 0, msp = 0;
    std::vector<TileDesc> tiles;
    std::vector<int> tile_of;                  // owned node (internal id) -> tile
    std::vector<int> halo_nodes, halo_elems, slot_elem;
    std::vector<unsigned long long> slot_conn; // 3 x 16-bit tile-local node ids
    std::vector<uint16_t> inc;                 // slot*3 + vertex, 0xFFFF = padding
    // state-resident solver only (resident_order = true)
    std::vector<ResTile> res_tiles;
    std::vector<int> res_nbr;                  // neighbour tile ids
    std::vector<uint16_t> res_n2n;             // [n2n_off + c*n_own + j] tile-local node id of the c-th neighbour (bamg order)
    std::vector<uint8_t> res_n2n_deg;          // [node_begin + j] = neighbour count
    std::vector<uint8_t> halo_move;            // per halo_nodes entry: this tile moves that ghost node (UM / UT)
    std::vector<int> halo_slot;                // per halo_nodes entry: mailbox slot the value is read from
    int n_export = 0;                          // export nodes of the rank; mailbox = [export nodes | ghost nodes]
    long p2_wavefronts = 0, p2_cells = 0;      // shared-memory wavefronts of one 64-bit gather of the nodal solve / ideal
};

// throws std::invalid_argument on inconsistent input
// resident_order: export-first node order, early-first slot order and the ResTile tables; export_mask (may be NULL)
// flags reference-numbered owned nodes that are sent to another rank; max_tile_nodes bounds the tile sizes when the cuts
// are moved to balance the slot counts.
void build_mesh_plan(const NsxMesh* M, MeshPlan& P, int target_tile_nodes, int sm_count,
                     bool resident_order = false, const uint8_t* export_mask = nullptr, int max_tile_nodes = 0);

}  // namespace nsx

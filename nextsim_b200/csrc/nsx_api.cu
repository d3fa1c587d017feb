// nsx_api.cu -- C ABI (include/nsx.h) over the CUDA kernels: handle life cycle, transfers, the
// explicitSolve()/update() launch sequences and the NVLink halo wiring.  No CPU fallback: every entry
// point that computes launches CUDA kernels and returns an error if the device is unavailable.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "nsx_kernels.cuh"

using namespace nsx;

static std::string g_create_err;

#define NSX_API_BEGIN(h)                                   \
    if (!(h)) return 1;                                    \
    try {                                                  \
        NSX_CUDA(cudaSetDevice((h)->device));
#define NSX_API_END(h)                                     \
    }                                                      \
    catch (std::exception const& e) {                      \
        (h)->err = e.what();                               \
        return 2;                                          \
    }                                                      \
    return 0;

static inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

extern "C" void nsx_create_options_defaults(NsxCreateOptions* o)
{
    if (!o) return;
    *o = NsxCreateOptions{};
    o->path = NSX_PATH_AUTO; o->tile_nodes = 0; o->max_sms = 0; o->use_graph = 1; o->overlap = 1; o->boundary_sms = 0; o->ow_skip = 1;
}

// ---------------------------------------------------------------------------------------------------
// shared-memory layout of the sub-cycle kernel for `n_node_planes` staged node planes
// ---------------------------------------------------------------------------------------------------
static SmemLayout sub_layout(MeshPlan const& P, int n_node_planes)
{
    auto al = [](size_t x) { return (x + 15) & ~size_t(15); };
    SmemLayout L{};
    size_t o = 0;
    L.msp = P.msp;
    L.mop = (P.max_own_slots + 3) & ~1;
    L.mtp = (P.tile_nodes + 3) & ~1;
    L.bar = (int)o;   o += 64;              // the tile descriptor travels with the stage
    L.conn = (int)o;  o += al((size_t)L.msp * 8);
    L.shape = (int)o; o += (size_t)6 * L.msp * 8;
    L.ec = (int)o;    o += (size_t)6 * L.msp * 8;
    L.sig = (int)o;   o += (size_t)3 * L.mop * 8;
    L.dmg = (int)o;   o += (size_t)L.mop * 8;
    L.node = (int)o;  o += (size_t)n_node_planes * L.mtp * 8;
    L.su = (int)o;    o += al(((size_t)P.max_local_nodes + 2) * 8);
    L.sv = (int)o;    o += al(((size_t)P.max_local_nodes + 2) * 8);
    L.mhs = (P.max_halo_slots + 3) & ~1;
    L.hsig = (int)o;  o += (size_t)4 * L.mhs * 8;
    L.inc = (int)o;   o += al((size_t)P.max_inc * 2 + 32);
    L.fl = (int)o;    o += al((size_t)P.tile_nodes + 32);
    L.total = (int)o;
    return L;
}
// dynamic shared memory of one k_resident CTA (its static part holds the tile descriptors; 1 KB per CTA is reserved by the driver)
constexpr int RES_SMEM_MAX = (227 * 1024 - (RES_CTAS - 1) * 1024) / RES_CTAS - 256;
constexpr int SUB_SMEM_CAP = ((227 * 1024) / SUB_CTAS_PER_SM - 1024 * (SUB_CTAS_PER_SM - 1) - 128) / SUB_STAGES;     // per stage; one persistent CTA per SM (227 KB)

// ---------------------------------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------------------------------
static void upload_plan(nsx_solver* S)
{
    MeshPlan& P = S->plan;
    cudaStream_t st = S->stream;
    S->nn = P.nn; S->ndof = P.ndof; S->ne = P.ne; S->ne_local = P.ne_local;
    S->node_perm.upload(P.node_perm, st); S->elem_perm.upload(P.elem_perm, st);
    S->x.upload(P.x, st); S->y.upload(P.y, st); S->lat.upload(P.lat, st);
    S->nflags.upload(P.nflags, st);
    S->en0.upload(P.en[0], st); S->en1.upload(P.en[1], st); S->en2.upload(P.en[2], st);
    S->n2e.upload(P.n2e, st); S->n2e_deg.upload(P.n2e_deg, st);
    S->nec.upload(P.nec, st);
    S->n2n.upload(P.n2n, st); S->n2n_deg.upload(P.n2n_deg, st);
    S->tiles.upload(P.tiles, st);
    S->halo_nodes.upload(P.halo_nodes, st); S->halo_elems.upload(P.halo_elems, st);
    S->slot_elem.upload(P.slot_elem, st); S->slot_conn.upload(P.slot_conn, st);
    S->inc.upload(P.inc, st);
    S->sub_smem = sub_layout(P, NP_COUNT).total;
    // the attribute is per function and device, shared by every handle: always ask for the cap
    NSX_CUDA(cudaFuncSetAttribute(k_subcycle<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    NSX_CUDA(cudaFuncSetAttribute(k_subcycle<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (S->resident) {
        S->res_tiles.upload(P.res_tiles, st); S->halo_slot.upload(P.halo_slot, st);
        S->res_n2n.upload(P.res_n2n, st); S->res_n2n_deg.upload(P.res_n2n_deg, st); S->halo_move.upload(P.halo_move, st);
        NSX_CUDA(cudaFuncSetAttribute(k_resident<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM_MAX));
        NSX_CUDA(cudaFuncSetAttribute(k_resident<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM_MAX));
    }
    NSX_CUDA(cudaStreamSynchronize(st));
}

static void alloc_fields(nsx_solver* S)
{
    size_t const nn = S->nn, ne = S->ne, ns = S->plan.nslots;
    cudaStream_t st = S->stream;
    // halo window (one allocation, so it can be exported over CUDA IPC): [VT0 | VT1 | flags | mailbox buffers 0, 1, 2]
    // mailbox (resident path): one 32-byte {u, tag, v, tag} entry per export node, then per ghost node
    size_t const vt_bytes = 2 * nn * sizeof(double);
    size_t const flag_bytes = 256 * sizeof(unsigned long long);
    S->n_mb = (S->resident ? S->plan.n_export : 0) + (S->nn - S->ndof);     // [export nodes (resident path) | ghost nodes]
    S->window_bytes = 2 * vt_bytes + flag_bytes + 3 * (size_t)S->n_mb * sizeof(MbEntry);     // 3 buffers (tile / direct paths; the resident path uses 2)
    NSX_CUDA(cudaMalloc(&S->window, S->window_bytes));
    NSX_CUDA(cudaMemsetAsync(S->window, 0, S->window_bytes, st));
    S->VT[0] = (double*)S->window;
    S->VT[1] = S->VT[0] + 2 * nn;
    S->flags = (unsigned long long*)(S->VT[1] + 2 * nn);
    S->mailbox = (void*)(S->flags + 256);

    DBuf<double>* nodal2[] = {&S->UM, &S->UT, &S->wind, &S->ocean, &S->tau_wi, &S->tau_a, &S->tau_w, &S->VTM, &S->grad_ssh};
    for (auto* b : nodal2) { b->alloc(2 * nn); b->zero(st); }
    DBuf<double>* nodal1[] = {&S->ssh, &S->node_mass, &S->rlmass, &S->cbu, &S->fcor};
    for (auto* b : nodal1) { b->alloc(nn); b->zero(st); }
    DBuf<double>* elem[] = {&S->sig[0][0], &S->sig[0][1], &S->sig[0][2], &S->sig[1][0], &S->sig[1][1], &S->sig[1][2],
                            &S->dmg[0], &S->dmg[1], &S->conc, &S->thick, &S->snow, &S->conc_young,
                            &S->h_young, &S->hs_young, &S->thick_myi, &S->conc_myi, &S->ridge_ratio, &S->depth,
                            &S->drag_ui, &S->drag_ui_young, &S->cohesion, &S->t_heal, &S->surface, &S->delta_x,
                            &S->del_ci_ridge_myi, &S->emass, &S->ecbu};
    for (auto* b : elem) { b->alloc(ne); b->zero(st); }
    S->shape.alloc(6 * ne); S->shape.zero(st);
    S->slot_shape.alloc(4 * ns); S->slot_shape.zero(st);
    S->disp.alloc(2 * nn); S->disp.zero(st);
    S->slot_ec.alloc(6 * ns); S->slot_ec.zero(st);
    NSX_CUDA(cudaMallocHost(&S->h_err, sizeof(int)));
    *S->h_err = 0;
    NSX_CUDA(cudaMallocHost(&S->h_time, 3 * sizeof(unsigned long long)));
    S->res_time.alloc(3); S->res_time.zero(st);
    // Path selection (NsxCreateOptions.path, AUTO by size): the sub-cycle state is resident in shared memory
    // (k_resident, decided in nsx_create_ex), or the working set (~300 B per element) lives in the 126 MB L2 (direct
    // path), or it streams from HBM (TMA tile pipeline).
    {
        S->direct = (double)ne * 300.0 < 90e6;
        if (S->opt.path == NSX_PATH_DIRECT) S->direct = true;
        if (S->opt.path == NSX_PATH_TILES) S->direct = false;
        if (S->resident) S->direct = false;
        if (S->direct) { S->ec_e.alloc(6 * ne); S->ec_e.zero(st); S->contrib.alloc(6 * ne); S->contrib.zero(st); }
    }
    S->ow_list.alloc(S->ndof); S->ow_count.alloc(1); S->ow_count.zero(st);
    S->check_i.alloc(4); S->check_d.alloc(1);
    S->halo_err.alloc(1); S->halo_err.zero(st);
    S->d_epoch.alloc(1); S->d_done.alloc(1);
    S->d_epoch.zero(st); S->d_done.zero(st);
    S->ow_pair.alloc(64); S->ow_pair.zero(st);
    for (auto& e : S->ev) NSX_CUDA(cudaEventCreate(&e));
    NSX_CUDA(cudaEventCreate(&S->ev_upd));
    NSX_CUDA(cudaEventCreateWithFlags(&S->ev_fork, cudaEventDisableTiming));
    NSX_CUDA(cudaEventCreateWithFlags(&S->ev_join, cudaEventDisableTiming));
}

static void build_halo(nsx_solver* S, const NsxHalo* H)
{
    S->rank = 0; S->nranks = 1;
    std::vector<int> order(S->plan.ntiles);
    for (int t = 0; t < S->plan.ntiles; ++t) order[t] = t;
    if (H) {
        S->rank = H->rank; S->nranks = H->nranks;
        if (H->nranks > 256) throw std::invalid_argument("nsx_create: at most 256 ranks");
        auto link = [&](int r) -> PeerLink& {
            for (auto& p : S->peers) if (p.rank == r) return p;
            S->peers.emplace_back();
            S->peers.back().rank = r;
            return S->peers.back();
        };
        std::vector<int> const& perm = S->plan.node_perm;
        for (int k = 0; k < H->n_send_peers; ++k) {
            PeerLink& p = link(H->send_peer[k]);
            for (int q = H->send_ptr[k]; q < H->send_ptr[k + 1]; ++q) {
                int const v = H->send_idx[q];
                if (v < 0 || v >= S->ndof) throw std::invalid_argument("nsx_create: send index is not an owned node");
                p.h_send_idx.push_back(perm[v]);
                S->plan.tiles[S->plan.tile_of[perm[v]]].boundary = 1;
            }
        }
        for (int k = 0; k < H->n_recv_peers; ++k) {
            PeerLink& p = link(H->recv_peer[k]);
            for (int q = H->recv_ptr[k]; q < H->recv_ptr[k + 1]; ++q) {
                int const v = H->recv_idx[q];
                if (v < S->ndof || v >= S->nn) throw std::invalid_argument("nsx_create: recv index is not a ghost node");
                p.h_recv_idx.push_back(perm[v]);
            }
        }
        if (S->peers.size() > 32) throw std::invalid_argument("nsx_create: at most 32 neighbour ranks");
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return S->plan.tiles[a].boundary > S->plan.tiles[b].boundary; });
    }
    // Every reader of a ghost slot must run BEFORE this rank publishes its epoch (a faster peer may overwrite the
    // slot for the sub-cycle after next as soon as it sees the flag): tiles that read ghost nodes are boundary
    // tiles too (flagged by build_mesh_plan), and the lagged ghost mesh move is spread over boundary tiles only.
    S->n_boundary_tiles = 0;
    for (auto const& td : S->plan.tiles) S->n_boundary_tiles += td.boundary;
    int const nghost = S->nn - S->ndof;
    if (nghost > 0) {
        int const nb = std::max(1, S->n_boundary_tiles);
        int const G = (nghost + nb - 1) / nb;
        int k = 0;
        for (int t : order) {
            nsx::TileDesc& td = S->plan.tiles[t];
            bool const take = (S->n_boundary_tiles == 0) ? (k == 0 && t == order[0]) : (td.boundary != 0);
            if (take) {
                td.ghost_begin = S->ndof + std::min(nghost, k * G);
                td.n_ghost = std::min(nghost, (k + 1) * G) - std::min(nghost, k * G);
                ++k;
            } else {
                td.ghost_begin = S->ndof; td.n_ghost = 0;
            }
        }
        S->tiles.upload(S->plan.tiles, S->stream);
    }
    S->tile_order.upload(order, S->stream);
    NSX_CUDA(cudaStreamSynchronize(S->stream));
}

extern "C" int nsx_version(void) { return NSX_VERSION; }

// Host-only checks of what the caller hands to nsx_create (no CUDA call): sizes, index ranges, list ordering.  The
// reference would fail later and less clearly (out-of-range vector access); here the handle is never created.
static void validate_inputs(const NsxMesh* M, const NsxHalo* H)
{
    auto fail = [](std::string const& m) { throw std::invalid_argument("nsx_create: " + m); };
    if (M->num_nodes <= 0 || M->num_elements <= 0) fail("the rank holds no nodes or no elements");
    if (M->local_ndof < 0 || M->local_ndof > M->num_nodes) fail("local_ndof outside [0, num_nodes]");
    if (M->local_nelements < 0 || M->local_nelements > M->num_elements) fail("local_nelements outside [0, num_elements]");
    if (!M->coord_x || !M->coord_y || !M->indices || !M->mask_dirichlet || !M->nodal_element_connectivity ||
        !M->nodal_connectivity || !M->lat)
        fail("NULL mesh array");
    if (M->n_neumann_flags < 0 || (M->n_neumann_flags > 0 && !M->neumann_flags)) fail("bad neumann_flags");
    if (M->nec_width < 1 || M->nc_width < 2) fail("connectivity table widths must be >= 1 (elements) and >= 2 (nodes)");
    for (long k = 0; k < 3L * M->num_elements; ++k)
        if (M->indices[k] < 1 || M->indices[k] > M->num_nodes)
            fail("indices[" + std::to_string(k) + "] = " + std::to_string(M->indices[k]) + " is not a 1-based local node id");
    for (int k = 0; k < M->n_neumann_flags; ++k) {
        if (M->neumann_flags[k] < 0 || M->neumann_flags[k] >= M->num_nodes) fail("neumann flag outside the local nodes");
        if (k && M->neumann_flags[k] <= M->neumann_flags[k - 1]) fail("neumann_flags must be sorted and unique (std::binary_search, FE.cpp:3957)");
    }
    for (int n = 0; n < M->num_nodes; ++n) {
        double const cnt = M->nodal_connectivity[(size_t)n * M->nc_width + M->nc_width - 1];
        if (!(cnt >= 0 && cnt <= M->nc_width - 1)) fail("NodalConnectivity count column out of range at node " + std::to_string(n));
    }
    if (M->local_ndof < M->num_nodes && !H) fail("ghost nodes without halo lists");
    if (!H) return;
    if (H->nranks < 1 || H->rank < 0 || H->rank >= H->nranks) fail("rank outside [0, nranks)");
    if (H->n_send_peers < 0 || H->n_recv_peers < 0) fail("negative peer count");
    if ((H->n_send_peers && (!H->send_peer || !H->send_ptr || !H->send_idx)) ||
        (H->n_recv_peers && (!H->recv_peer || !H->recv_ptr || !H->recv_idx)))
        fail("NULL halo list");
    for (int k = 0; k < H->n_send_peers; ++k) {
        if (H->send_peer[k] < 0 || H->send_peer[k] >= H->nranks || H->send_peer[k] == H->rank) fail("bad send peer");
        if (H->send_ptr[k + 1] < H->send_ptr[k]) fail("send_ptr not monotone");
    }
    for (int k = 0; k < H->n_recv_peers; ++k) {
        if (H->recv_peer[k] < 0 || H->recv_peer[k] >= H->nranks || H->recv_peer[k] == H->rank) fail("bad receive peer");
        if (H->recv_ptr[k + 1] < H->recv_ptr[k]) fail("recv_ptr not monotone");
    }
    int nrecv = H->n_recv_peers ? H->recv_ptr[H->n_recv_peers] : 0;
    if (nrecv != M->num_nodes - M->local_ndof) fail("the receive lists must cover every ghost node exactly once (" +
        std::to_string(nrecv) + " entries for " + std::to_string(M->num_nodes - M->local_ndof) + " ghosts)");
}

// exported so that hosts (and the CPU test-suite) can check a mesh without a device; same messages as nsx_create
extern "C" int nsx_validate_mesh(const NsxMesh* mesh, const NsxHalo* halo)
{
    try {
        if (!mesh) throw std::invalid_argument("nsx_create: NULL argument");
        validate_inputs(mesh, halo);
    } catch (std::exception const& e) { g_create_err = e.what(); return 2; }
    return 0;
}

// The state-resident plan: one tile per CTA, at most `ctas` tiles, everything the sub-cycle loop touches in shared
// memory and registers.  Returns false (with the reason) when the mesh does not fit; AUTO then falls back.
static bool try_resident_plan(const NsxMesh* mesh, const NsxHalo* halo, int ctas, MeshPlan& P, std::string& why)
{
    std::vector<uint8_t> xmask;
    int nlinks = 0;
    if (halo) {
        xmask.assign(mesh->num_nodes, 0);
        int const ns = halo->n_send_peers ? halo->send_ptr[halo->n_send_peers] : 0;
        for (int q = 0; q < ns; ++q) {
            int const v = halo->send_idx[q];
            if (v < 0 || v >= mesh->local_ndof) { why = "send index is not an owned node"; return false; }
            xmask[v] = 1;
        }
        std::vector<int> peers(halo->send_peer, halo->send_peer + halo->n_send_peers);
        peers.insert(peers.end(), halo->recv_peer, halo->recv_peer + halo->n_recv_peers);
        std::sort(peers.begin(), peers.end());
        nlinks = (int)(std::unique(peers.begin(), peers.end()) - peers.begin());
    }
    if (nlinks > RES_MAX_LINKS) { why = "more than " + std::to_string(RES_MAX_LINKS) + " neighbour ranks"; return false; }
    int const T = (mesh->local_ndof + ctas - 1) / ctas;
    if (T > RES_TPB) { why = std::to_string(T) + " owned nodes per tile (limit " + std::to_string(RES_TPB) + ")"; return false; }
    P = MeshPlan();
    build_mesh_plan(mesh, P, std::max(32, T), 1, true, halo ? xmask.data() : nullptr, RES_TPB);
    size_t const smem = (size_t)(16 * P.msp + 2 * (P.max_local_nodes + 2)) * sizeof(double);     // BBM: 6 + 6 + 3 + 1 planes
    if (P.ntiles > ctas || P.tile_nodes > RES_TPB || P.max_slots > RES_SPT * RES_TPB || smem > (size_t)RES_SMEM_MAX || P.msp >= 16384) {
        why = "tiles " + std::to_string(P.ntiles) + ", nodes per tile " + std::to_string(P.tile_nodes) + ", max slots " +
              std::to_string(P.max_slots) + ", shared memory " + std::to_string(smem) + " B";
        return false;
    }
    // k_resident updates one late slot per thread before the halo wait and at most two slots per thread after it
    for (size_t t = 0; t < P.tiles.size() && t < P.res_tiles.size(); ++t) {
        int const E = P.res_tiles[t].n_early_own, O = P.tiles[t].n_own_slots, nsl = O + P.tiles[t].n_halo_slots;
        int const after = E + nsl - std::min(O, E + RES_TPB);
        if (after > (RES_SPT - 1) * RES_TPB) {
            why = "tile " + std::to_string(t) + ": " + std::to_string(after) + " slots after the halo wait (limit " +
                  std::to_string((RES_SPT - 1) * RES_TPB) + ")";
            return false;
        }
    }
    return true;
}

extern "C" int nsx_create(const NsxMesh* mesh, const NsxHalo* halo, int device, nsx_handle* out)
{
    return nsx_create_ex(mesh, halo, device, nullptr, out);
}

extern "C" int nsx_device_sm_count(int device)
{
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    return n;
}

extern "C" int nsx_create_ex(const NsxMesh* mesh, const NsxHalo* halo, int device, const NsxCreateOptions* opt, nsx_handle* out)
{
    if (!mesh || !out) { g_create_err = "nsx_create: NULL argument"; return 1; }
    *out = nullptr;
    nsx_solver* S = new nsx_solver();
    try {
        nsx_create_options_defaults(&S->opt);
        if (opt) S->opt = *opt;
        if (S->opt.path < NSX_PATH_AUTO || S->opt.path > NSX_PATH_RESIDENT) throw std::invalid_argument("nsx_create: unknown path option");
        validate_inputs(mesh, halo);
        int ndev = 0;
        NSX_CUDA(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev) throw std::invalid_argument("nsx_create: no such CUDA device");
        S->device = device;
        NSX_CUDA(cudaSetDevice(device));
        NSX_CUDA(cudaDeviceGetAttribute(&S->sm_count, cudaDevAttrMultiProcessorCount, device));
        NSX_CUDA(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));
        NSX_CUDA(cudaStreamCreateWithFlags(&S->stream2, cudaStreamNonBlocking));
        int const ctas = RES_CTAS * ((S->opt.max_sms > 0) ? std::min(S->opt.max_sms, S->sm_count) : S->sm_count);
        // resident path: explicit request, or AUTO when the mesh fits
        S->resident = false;
        if (S->opt.path == NSX_PATH_RESIDENT || S->opt.path == NSX_PATH_AUTO) {
            std::string why;
            S->resident = try_resident_plan(mesh, halo, ctas, S->plan, why);
            if (!S->resident && S->opt.path == NSX_PATH_RESIDENT)
                throw std::invalid_argument("nsx_create: the resident path does not fit this mesh (" + why + ")");
        }
        if (!S->resident) {
            // one wave of the sub-cycle kernel = SMs x resident CTAs; shrink the tiles until the staged working set of
            // the largest tile fits the per-CTA shared-memory budget
            int target = S->opt.tile_nodes > 0 ? S->opt.tile_nodes : 208;
            for (int attempt = 0;; ++attempt) {
                S->plan = MeshPlan();
                build_mesh_plan(mesh, S->plan, target, S->sm_count * SUB_CTAS_PER_SM);
                if (sub_layout(S->plan, 14).total <= SUB_SMEM_CAP) break;
                if (attempt > 12 || target <= 32) throw std::invalid_argument("nsx_create: cannot fit a tile in shared memory");
                target = std::max(32, (int)(target * 0.88));
            }
        }
        upload_plan(S);
        alloc_fields(S);
        build_halo(S, halo);
        NSX_CUDA(cudaStreamSynchronize(S->stream));
    } catch (std::exception const& e) {
        g_create_err = e.what();
        nsx_destroy(S);
        return 2;
    }
    *out = S;
    return 0;
}

extern "C" int nsx_destroy(nsx_handle S)
{
    if (!S) return 0;
    cudaSetDevice(S->device);
    if (S->stream) cudaStreamSynchronize(S->stream);
    if (S->stream2) cudaStreamSynchronize(S->stream2);
    for (auto& p : S->peers) if (p.ipc_base) cudaIpcCloseMemHandle(p.ipc_base);
    for (auto& g : S->graph_exec) if (g) cudaGraphExecDestroy(g);
    for (auto& e : S->ev) if (e) cudaEventDestroy(e);
    if (S->ev_upd) cudaEventDestroy(S->ev_upd);
    if (S->ev_fork) cudaEventDestroy(S->ev_fork);
    if (S->ev_join) cudaEventDestroy(S->ev_join);
    if (S->window) cudaFree(S->window);
    if (S->h_err) cudaFreeHost(S->h_err);
    if (S->h_time) cudaFreeHost(S->h_time);
    if (S->stream2) cudaStreamDestroy(S->stream2);
    if (S->stream) cudaStreamDestroy(S->stream);
    delete S;
    return 0;
}

extern "C" const char* nsx_last_error(nsx_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }
extern "C" void* nsx_get_stream(nsx_handle h) { return h ? (void*)h->stream : nullptr; }

// sizeof() of every struct that crosses the ABI, so a binding can detect layout drift
extern "C" int nsx_abi_sizes(int* out, int n)
{
    int const s[9] = {(int)sizeof(NsxDynParams), (int)sizeof(NsxMesh), (int)sizeof(NsxHalo),
                      (int)sizeof(NsxFields), (int)sizeof(NsxCheck), (int)sizeof(NsxTiming), (int)sizeof(NsxRegrid),
                      (int)sizeof(NsxCreateOptions), (int)sizeof(NsxThermoParams)};
    for (int i = 0; i < n && i < 9; ++i) out[i] = s[i];
    return 9;
}

// tile decomposition summary: ntiles, nodes per tile, slots, max local nodes, max slots, boundary tiles, smem bytes,
// and whether the direct (L2-resident) path is selected
extern "C" int nsx_tile_info(nsx_handle S, int* out, int n)
{
    if (!S) return 0;
    int const v[8] = {S->plan.ntiles, S->plan.tile_nodes, S->plan.nslots, S->plan.max_local_nodes, S->plan.max_slots,
                      S->n_boundary_tiles, (int)S->sub_smem,
                      S->resident ? NSX_PATH_RESIDENT : S->direct ? NSX_PATH_DIRECT : NSX_PATH_TILES};
    for (int i = 0; i < n && i < 8; ++i) out[i] = v[i];
    return 8;
}

// ---------------------------------------------------------------------------------------------------
// options -> kernel scalars
// ---------------------------------------------------------------------------------------------------
static KParams derive(nsx_solver const* S, NsxDynParams const& P)
{
    KParams K{};
    K.dynamics_type = P.dynamics_type;
    K.basal_stress_type = P.basal_stress_type;
    K.young_ice = (P.ice_cat_type == NSX_ICECAT_YOUNG_ICE);
    K.nn = S->nn; K.ndof = S->ndof; K.ne = S->ne;
    K.dtime_step = P.dtime_step;
    K.dte = P.dtime_step / double(P.substeps);                 // FE.cpp:10185
    K.cos_ota = std::cos(P.ocean_turning_angle_rad);           // FE.cpp:10187-10188
    K.sin_ota_abs = std::fabs(std::sin(P.ocean_turning_angle_rad));
    K.min_m = RHOI * P.min_h;                                  // FE.cpp:10191
    K.young = P.young;
    K.compaction_param = P.compaction_param;
    K.lambda0 = P.undamaged_time_relaxation_sigma;
    K.exp_relax_m1 = P.exponent_relaxation_sigma - 1.;
    K.relax_int_pow = -1;
    for (int k = 0; k <= 4; ++k) if (K.exp_relax_m1 == (double)k) K.relax_int_pow = k;
    K.compression_factor = P.compression_factor;
    K.exp_compression = P.exponent_compression_factor;
    K.compr_strength = P.compr_strength;
    K.tan_phi = P.tan_phi;
    K.sqrt_nu_rhoi = std::sqrt(2. * (1. + P.nu0) * RHOI);      // FE.cpp:4140
    double const Dunit_factor = 1. / (1. - P.nu0 * P.nu0);     // FE.cpp:1495-1505
    K.D00 = Dunit_factor * 1.;
    K.D01 = Dunit_factor * P.nu0;
    K.D22 = Dunit_factor * (1. - P.nu0) / 2.;
    K.evp_e = P.evp_e; K.evp_Pstar = P.evp_Pstar; K.evp_C = P.evp_C; K.evp_dmin = P.evp_dmin;
    K.re2 = 1. / (P.evp_e * P.evp_e);
    if (P.dynamics_type == NSX_DYN_EVP) {                      // FE.cpp:10708-10711
        double const T = P.dtime_step / 3.;
        K.ralpha1 = 0.5 * K.dte / T;
        K.ralpha2 = 0.5 * K.dte / T * P.evp_e * P.evp_e;
    } else {                                                   // FE.cpp:10724
        K.ralpha1 = 1. / P.mevp_alpha;
        K.ralpha2 = 1. / P.mevp_alpha;
    }
    K.mevp_b = P.mevp_beta + 1.;
    K.mevp_rb = 1. / K.mevp_b;
    K.dte_mevp = K.dte / K.mevp_b;
    K.rhow_cdw = RHOW * P.quad_drag_coef_water;
    K.u0 = P.basal_u0;
    K.k1 = P.basal_k1; K.k2 = P.basal_k2; K.Cb = P.basal_Cb;
    K.min_c = P.min_c; K.min_h = P.min_h;
    K.equal_ridging = P.equal_ridging;
    K.myi_with_young = (P.newice_type == 4 && P.use_young_ice_in_myi_reset);
    return K;
}

extern "C" int nsx_set_params(nsx_handle S, const NsxDynParams* p)
{
    NSX_API_BEGIN(S)
    if (!p) throw std::invalid_argument("nsx_set_params: NULL");
    if (p->dynamics_type != NSX_DYN_BBM && p->dynamics_type != NSX_DYN_EVP && p->dynamics_type != NSX_DYN_MEVP)
        throw std::invalid_argument("nsx_set_params: dynamics_type must be bbm, evp or mevp");
    if (p->substeps <= 0 || !(p->dtime_step > 0.)) throw std::invalid_argument("nsx_set_params: substeps and dtime_step must be positive");
    S->P = *p;
    S->K = derive(S, *p);
    S->have_params = true;
    S->graph_valid = false;
    NSX_API_END(S)
}

// ---------------------------------------------------------------------------------------------------
// transfers: the host speaks the reference's local numbering, the device an internal one; every field goes
// through a device staging buffer and a permutation kernel (stream ordered, no extra synchronisation).
// ---------------------------------------------------------------------------------------------------
namespace {
enum Kind { NODAL2, NODAL1, ELEM };
struct FieldMap { double* host; double* dev; Kind kind; };
}
static void field_table(nsx_solver* S, const NsxFields* f, std::vector<FieldMap>& t)
{
    auto add = [&](double* h, double* d, Kind k) { if (h) t.push_back({h, d, k}); };
    add(f->M_VT, S->VT[S->cur], NODAL2);
    add(f->M_UM, S->UM.p, NODAL2);
    add(f->M_UT, S->UT.p, NODAL2);
    add(f->M_wind, S->wind.p, NODAL2);
    add(f->M_ocean, S->ocean.p, NODAL2);
    add(f->M_tau_wi, S->tau_wi.p, NODAL2);
    add(f->D_tau_a, S->tau_a.p, NODAL2);
    add(f->D_tau_w, S->tau_w.p, NODAL2);
    add(f->M_ssh, S->ssh.p, NODAL1);
    add(f->M_sigma[0], S->sig[S->scur][0].p, ELEM);
    add(f->M_sigma[1], S->sig[S->scur][1].p, ELEM);
    add(f->M_sigma[2], S->sig[S->scur][2].p, ELEM);
    add(f->M_damage, S->dmg[S->dcur].p, ELEM);
    add(f->M_conc, S->conc.p, ELEM);
    add(f->M_thick, S->thick.p, ELEM);
    add(f->M_snow_thick, S->snow.p, ELEM);
    add(f->M_conc_young, S->conc_young.p, ELEM);
    add(f->M_h_young, S->h_young.p, ELEM);
    add(f->M_hs_young, S->hs_young.p, ELEM);
    add(f->M_thick_myi, S->thick_myi.p, ELEM);
    add(f->M_conc_myi, S->conc_myi.p, ELEM);
    add(f->M_ridge_ratio, S->ridge_ratio.p, ELEM);
    add(f->M_element_depth, S->depth.p, ELEM);
    add(f->M_drag_ui, S->drag_ui.p, ELEM);
    add(f->M_drag_ui_young, S->drag_ui_young.p, ELEM);
    add(f->M_Cohesion, S->cohesion.p, ELEM);
    add(f->M_time_relaxation_damage, S->t_heal.p, ELEM);
    add(f->M_surface, S->surface.p, ELEM);
    add(f->M_delta_x, S->delta_x.p, ELEM);
    add(f->D_del_ci_ridge_myi, S->del_ci_ridge_myi.p, ELEM);
    double* const dg[6] = {f->D_conc, f->D_thick, f->D_snow_thick, f->D_sigma[0], f->D_sigma[1], f->D_divergence};
    for (int k = 0; k < 6; ++k) {
        if (!dg[k]) continue;
        if (!S->diag.p) throw std::runtime_error("diagnostics requested before nsx_update_ice_diagnostics");
        add(dg[k], S->diag.p + (size_t)k * S->ne, ELEM);
    }
}

// device error word written by the bounded spins of the exchange / barrier kernels
static std::string halo_error_text(int herr)
{
    if (herr >= 3000) return "device-side bounds check " + std::to_string(herr - 3000) + " failed (NSX_DEBUG_CHECKS build, nsx_kernels.cuh)";
    if (herr >= 2000) return "open-water smoother: grid barrier timed out (the launch was not co-resident)";
    if (herr >= 1000) return "resident solver: timed out waiting for the flag of tile " + std::to_string(herr - 1000) +
                             " (the launch was not co-resident, or a neighbour rank died)";
    return "halo exchange timed out waiting for rank " + std::to_string(herr - 1);
}

// One call = one device arena in host order: the PCIe copies of all fields are issued back to back on the handle's
// stream (one cudaMemcpyAsync per caller array -- the host keeps its own vectors, pinned or not), ONE kernel permutes
// every field between host and internal numbering, ONE stream synchronisation ends the call.
static size_t xfer_layout(nsx_solver* S, std::vector<FieldMap> const& t, std::vector<size_t>& off)
{
    size_t total = 0;
    off.resize(t.size());
    for (size_t k = 0; k < t.size(); ++k) {
        size_t const n = (t[k].kind == ELEM) ? S->ne : S->nn;
        off[k] = total;
        total += n * ((t[k].kind == NODAL2) ? 2 : 1);
    }
    if (S->arena.n < total) {
        NSX_CUDA(cudaStreamSynchronize(S->stream));
        S->arena.alloc(total + total / 8);
    }
    return total;
}

template <int IN>
static void xfer_permute(nsx_solver* S, std::vector<FieldMap> const& t, std::vector<size_t> const& off)
{
    for (size_t k0 = 0; k0 < t.size(); k0 += XFER_MAX) {
        XferTable T{};
        int nmax = 0;
        for (size_t k = k0; k < t.size() && k < k0 + XFER_MAX; ++k) {
            XferEnt& e = T.e[T.count++];
            e.n = (t[k].kind == ELEM) ? S->ne : S->nn;
            e.planes = (t[k].kind == NODAL2) ? 2 : 1;
            e.elem = (t[k].kind == ELEM) ? 1 : 0;
            if (IN) { e.src = S->arena.p + off[k]; e.dst = t[k].dev; }
            else { e.src = t[k].dev; e.dst = S->arena.p + off[k]; }
            nmax = std::max(nmax, e.n);
        }
        dim3 const grid(std::min(nblk(nmax), 4 * S->sm_count), T.count);
        k_permute_all<IN><<<grid, TPB, 0, S->stream>>>(T, S->node_perm.p, S->elem_perm.p);
    }
    NSX_CUDA(cudaGetLastError());
}

extern "C" int nsx_upload(nsx_handle S, const NsxFields* f)
{
    NSX_API_BEGIN(S)
    if (!f) throw std::invalid_argument("nsx_upload: NULL");
    if (f->M_shape_coeff) throw std::invalid_argument("nsx_upload: M_shape_coeff is an output");
    if (f->D_conc || f->D_thick || f->D_snow_thick || f->D_sigma[0] || f->D_sigma[1] || f->D_divergence)
        throw std::invalid_argument("nsx_upload: the D_* diagnostics are outputs");
    std::vector<FieldMap> t;
    field_table(S, f, t);
    std::vector<size_t> off;
    xfer_layout(S, t, off);
    cudaStream_t st = S->stream;
    for (size_t k = 0; k < t.size(); ++k) {
        size_t const n = (t[k].kind == ELEM) ? S->ne : S->nn;
        NSX_CUDA(cudaMemcpyAsync(S->arena.p + off[k], t[k].host, n * ((t[k].kind == NODAL2) ? 2 : 1) * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
    }
    if (!t.empty()) xfer_permute<1>(S, t, off);
    if (f->M_tau_wi && !S->have_tau_wi) { S->have_tau_wi = true; S->graph_valid = false; }
    NSX_CUDA(cudaStreamSynchronize(st));
    NSX_API_END(S)
}

extern "C" int nsx_download(nsx_handle S, NsxFields* f)
{
    NSX_API_BEGIN(S)
    if (!f) throw std::invalid_argument("nsx_download: NULL");
    std::vector<FieldMap> t;
    field_table(S, f, t);
    std::vector<size_t> off;
    size_t const total = xfer_layout(S, t, off);
    if (f->M_shape_coeff && S->arena.n < total + 6 * (size_t)S->ne) {      // the 6 planes go behind the other fields
        NSX_CUDA(cudaStreamSynchronize(S->stream));
        S->arena.alloc(total + 6 * (size_t)S->ne);
    }
    cudaStream_t st = S->stream;
    if (!t.empty()) xfer_permute<0>(S, t, off);
    for (size_t k = 0; k < t.size(); ++k) {
        size_t const n = (t[k].kind == ELEM) ? S->ne : S->nn;
        NSX_CUDA(cudaMemcpyAsync(t[k].host, S->arena.p + off[k], n * ((t[k].kind == NODAL2) ? 2 : 1) * sizeof(double),
                                 cudaMemcpyDeviceToHost, st));
    }
    if (f->M_shape_coeff) {
        k_shape_out<<<nblk(6L * S->ne), TPB, 0, st>>>(S->ne, S->elem_perm.p, S->shape.p, S->arena.p + total);
        NSX_CUDA(cudaMemcpyAsync(f->M_shape_coeff, S->arena.p + total, 6 * (size_t)S->ne * sizeof(double), cudaMemcpyDeviceToHost, st));
        NSX_CUDA(cudaGetLastError());
    }
    // the device error word (bounded spins of the exchange kernels) rides in the same stream, into pinned memory
    NSX_CUDA(cudaMemcpyAsync(S->h_err, S->halo_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    NSX_CUDA(cudaStreamSynchronize(st));
    if (*S->h_err) throw std::runtime_error(halo_error_text(*S->h_err));
    NSX_API_END(S)
}

extern "C" int nsx_host_register(void* p, unsigned long bytes)
{
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : 2;
}
extern "C" int nsx_host_unregister(void* p) { return cudaHostUnregister(p) == cudaSuccess ? 0 : 2; }

extern "C" int nsx_synchronize(nsx_handle S)
{
    NSX_API_BEGIN(S)
    NSX_CUDA(cudaStreamSynchronize(S->stream));
    NSX_API_END(S)
}

// ---------------------------------------------------------------------------------------------------
// halo wiring
// ---------------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == NSX_IPC_HANDLE_BYTES, "ipc handle size");

// the peer's window: [VT0 | VT1 | flags | mailbox]; hdr = {num_nodes, list length, local_ndof, export nodes, resident} of the peer
static void finish_link(nsx_solver* S, PeerLink& p, double* base, const int* hdr, const int* peer_recv_idx_for_me, size_t n)
{
    if (n != p.h_send_idx.size()) throw std::runtime_error("halo: peer ghost list length differs from my send list");
    int const peer_nn = hdr[0], peer_ndof = hdr[2], peer_nx = hdr[3], peer_resident = hdr[4];
    p.peer_nn = peer_nn;
    p.peer_vt[0] = base;
    p.peer_vt[1] = base + 2 * (size_t)peer_nn;
    p.peer_flags = (unsigned long long*)(base + 4 * (size_t)peer_nn);
    p.h_send_dst.assign(peer_recv_idx_for_me, peer_recv_idx_for_me + n);
    // my values go into the peer's mailbox: ghost g -> slot (its export nodes, resident path only) + (g - ndof)
    p.peer_mb = (void*)(p.peer_flags + 256);
    p.peer_nmb = peer_nx + (peer_nn - peer_ndof);
    p.h_send_slot.clear();
    for (size_t k = 0; k < n; ++k) p.h_send_slot.push_back(peer_nx + (peer_recv_idx_for_me[k] - peer_ndof));
    if (S->resident != (peer_resident != 0)) throw std::runtime_error("halo: neighbour ranks must use the same sub-cycle path");
    p.connected = true;
}

extern "C" int nsx_halo_connect_local(nsx_handle S, int peer_rank, nsx_handle Q)
{
    NSX_API_BEGIN(S)
    if (!Q) throw std::invalid_argument("nsx_halo_connect_local: NULL peer");
    PeerLink* mine = nullptr;
    for (auto& p : S->peers) if (p.rank == peer_rank) mine = &p;
    if (!mine) throw std::invalid_argument("nsx_halo_connect_local: not a neighbour rank");
    const PeerLink* theirs = nullptr;
    for (auto& p : Q->peers) if (p.rank == S->rank) theirs = &p;
    static const std::vector<int> empty;
    auto const& ridx = theirs ? theirs->h_recv_idx : empty;
    if (Q->device != S->device) {
        int can = 0;
        NSX_CUDA(cudaDeviceCanAccessPeer(&can, S->device, Q->device));
        if (!can) throw std::runtime_error("nsx_halo_connect_local: no peer access between the two devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(Q->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) NSX_CUDA(e);
        (void)cudaGetLastError();
    }
    int const hdr[5] = {Q->nn, (int)ridx.size(), Q->ndof, Q->resident ? Q->plan.n_export : 0, Q->resident ? 1 : 0};
    finish_link(S, *mine, Q->VT[0], hdr, ridx.data(), ridx.size());
    NSX_API_END(S)
}

// The peer's ghost index list for me travels with the IPC handle:
// [64 B handle | int num_nodes | int n | int local_ndof | int export nodes | int resident path | n ints]
constexpr int BLOB_HDR = 20;
extern "C" int nsx_halo_blob_size(nsx_handle S, int peer_rank)
{
    if (!S) return -1;
    for (auto& p : S->peers) if (p.rank == peer_rank) return NSX_IPC_HANDLE_BYTES + BLOB_HDR + 4 * (int)p.h_recv_idx.size();
    return NSX_IPC_HANDLE_BYTES + BLOB_HDR;
}
extern "C" int nsx_halo_blob(nsx_handle S, int peer_rank, unsigned char* out)
{
    NSX_API_BEGIN(S)
    NSX_CUDA(cudaIpcGetMemHandle(&S->ipc, S->window));
    std::memcpy(out, &S->ipc, NSX_IPC_HANDLE_BYTES);
    int hdr[5] = {S->nn, 0, S->ndof, S->resident ? S->plan.n_export : 0, S->resident ? 1 : 0};
    const PeerLink* pl = nullptr;
    for (auto& p : S->peers) if (p.rank == peer_rank) pl = &p;
    if (pl) hdr[1] = (int)pl->h_recv_idx.size();
    std::memcpy(out + NSX_IPC_HANDLE_BYTES, hdr, BLOB_HDR);
    if (pl && hdr[1]) std::memcpy(out + NSX_IPC_HANDLE_BYTES + BLOB_HDR, pl->h_recv_idx.data(), 4 * (size_t)hdr[1]);
    NSX_API_END(S)
}
extern "C" int nsx_halo_connect_blob(nsx_handle S, int peer_rank, const unsigned char* blob)
{
    NSX_API_BEGIN(S)
    PeerLink* mine = nullptr;
    for (auto& p : S->peers) if (p.rank == peer_rank) mine = &p;
    if (!mine) throw std::invalid_argument("nsx_halo_connect_blob: not a neighbour rank");
    cudaIpcMemHandle_t hdl;
    std::memcpy(&hdl, blob, NSX_IPC_HANDLE_BYTES);
    int hdr[5];
    std::memcpy(hdr, blob + NSX_IPC_HANDLE_BYTES, BLOB_HDR);
    void* base = nullptr;
    NSX_CUDA(cudaIpcOpenMemHandle(&base, hdl, cudaIpcMemLazyEnablePeerAccess));
    mine->ipc_base = base;
    finish_link(S, *mine, (double*)base, hdr, (const int*)(blob + NSX_IPC_HANDLE_BYTES + BLOB_HDR), (size_t)hdr[1]);
    NSX_API_END(S)
}

extern "C" int nsx_halo_finalize(nsx_handle S)
{
    NSX_API_BEGIN(S)
    for (auto& p : S->peers) if (!p.connected) throw std::runtime_error("nsx_halo_finalize: rank " + std::to_string(p.rank) + " not connected");
    // concatenated push tables (entry -> my node id, holder's ghost id), peers in link order
    std::vector<int> src, dst;
    for (auto& p : S->peers) {
        src.insert(src.end(), p.h_send_idx.begin(), p.h_send_idx.end());
        dst.insert(dst.end(), p.h_send_dst.begin(), p.h_send_dst.end());
    }
    S->n_send_total = (int)src.size();
    S->d_send_src.upload(src, S->stream); S->d_send_dst.upload(dst, S->stream);
    {   // per send entry: (send slot, slot in the holder's mailbox), same order as d_send_src
        std::vector<int2> es;
        int sl = 0;
        for (auto& p : S->peers) {
            if (p.h_send_idx.empty()) continue;
            for (size_t k = 0; k < p.h_send_idx.size(); ++k) es.push_back(make_int2(sl, p.h_send_slot[k]));
            ++sl;
        }
        S->d_send_slot.upload(es, S->stream);
    }
    // per-node push lists for the fused boundary launch: owned node -> (send-peer slot, holder's ghost id)
    {
        std::vector<int> ptr(S->ndof + 1, 0);
        int slot = 0;
        for (auto& p : S->peers) {
            if (p.h_send_idx.empty()) continue;
            for (int n : p.h_send_idx) ptr[n + 1]++;
            ++slot;
        }
        for (int n = 0; n < S->ndof; ++n) ptr[n + 1] += ptr[n];
        std::vector<int2> ent(ptr[S->ndof]);
        std::vector<int> fill(ptr.begin(), ptr.end() - 1);
        slot = 0;
        for (auto& p : S->peers) {
            if (p.h_send_idx.empty()) continue;
            for (size_t k = 0; k < p.h_send_idx.size(); ++k) ent[fill[p.h_send_idx[k]]++] = make_int2(slot, p.h_send_dst[k]);
            ++slot;
        }
        S->push_ptr.upload(ptr, S->stream);
        S->push_ent.upload(ent, S->stream);
        {                               // same lists with the holder's mailbox slot as destination
            std::vector<int2> entm(ent.size());
            std::vector<int> fillm(ptr.begin(), ptr.end() - 1);
            int sl = 0;
            for (auto& p : S->peers) {
                if (p.h_send_idx.empty()) continue;
                for (size_t k = 0; k < p.h_send_idx.size(); ++k) entm[fillm[p.h_send_idx[k]]++] = make_int2(sl, p.h_send_slot[k]);
                ++sl;
            }
            if (sl > RES_MAX_LINKS) throw std::runtime_error("resident path: too many neighbour ranks");
            S->push_ent_mb.upload(entm, S->stream);
        }
        // mixed direct/tile mode: elements written by boundary tiles, nodes owned by boundary tiles
        std::vector<uint8_t> nowrite(S->ne, 0), fl = S->plan.nflags;
        for (auto const& td : S->plan.tiles) {
            if (!td.boundary) continue;
            for (int k = 0; k < td.n_own_slots; ++k) nowrite[td.elem_begin + k] = 1;
            for (int j = 0; j < td.n_own; ++j) fl[td.node_begin + j] |= NF_BTILE;
        }
        S->elem_nowrite.upload(nowrite, S->stream);
        S->nflags.upload(fl, S->stream);
    }
    NSX_CUDA(cudaStreamSynchronize(S->stream));
    S->halo_ready = true;
    S->graph_valid = false;
    NSX_API_END(S)
}

// One ghost exchange of VT[cur] as a single kernel: push my owned values into every holder, publish the
// epoch, wait for every owner of my ghosts.  `sync` is false for lock-step groups ordered by streams/events.
static HaloArgs halo_args(nsx_solver* S, int parity, bool sync)
{
    HaloArgs a{};
    a.sync = sync ? 1 : 0;
    int off = 0;
    for (auto& p : S->peers) {
        if (!p.h_send_idx.empty()) {
            a.peer_begin[a.n_peers] = off;
            a.peer_vt[a.n_peers] = p.peer_vt[parity];
            a.peer_nn[a.n_peers] = p.peer_nn;
            a.peer_link[a.n_peers] = a.n_link;
            off += (int)p.h_send_idx.size();
            a.n_peers++;
        }
        a.link_flag[a.n_link] = p.peer_flags + S->rank;
        a.link_rank[a.n_link] = p.rank;
        a.n_link++;
    }
    a.peer_begin[a.n_peers] = off;
    a.n_total = off;
    return a;
}

static void halo_exchange(nsx_solver* S, bool sync)
{
    if (S->peers.empty()) return;
    if (!S->halo_ready) throw std::runtime_error("halo exchange before nsx_halo_finalize");
    HaloArgs a = halo_args(S, S->cur, sync);
    int const off = a.n_total;
    k_halo_exchange<<<std::max(1, nblk(off)), TPB, 0, S->stream>>>(a, S->nn, S->d_send_src.p, S->d_send_dst.p,
        S->VT[S->cur], S->flags, S->d_epoch.p, S->d_done.p, 40000000LL, S->halo_err.p);
    S->n_launch++;
    NSX_CUDA(cudaGetLastError());
}

extern "C" int nsx_update_ghosts(nsx_handle S)
{
    NSX_API_BEGIN(S)
    halo_exchange(S, !S->halo_local);
    NSX_API_END(S)
}

// ---------------------------------------------------------------------------------------------------
// explicitSolve()  FE.cpp:10182-10643 as a launch sequence (phases split so that several ranks living
// in one process can be stepped in lock-step on one device: nsx_group_explicit_solve)
// ---------------------------------------------------------------------------------------------------
static void phase_prep(nsx_solver* S)
{
    cudaStream_t st = S->stream;
    KParams const& K = S->K;
    S->ow_count.zero(st);
    k_prep_elements<<<nblk(S->plan.nslots), TPB, 0, st>>>(K, S->plan.nslots, S->slot_elem.p, S->en0.p, S->en1.p, S->en2.p,
        S->x.p, S->y.p, S->UM.p, S->conc.p, S->thick.p, S->snow.p, S->conc_young.p, S->h_young.p, S->hs_young.p,
        S->depth.p, S->ssh.p, S->cohesion.p, S->t_heal.p, S->surface.p, S->delta_x.p, S->shape.p, S->emass.p, S->ecbu.p,
        S->slot_shape.p, S->slot_ec.p, S->direct ? S->ec_e.p : nullptr);
    k_prep_nodes<<<nblk(S->nn), TPB, 0, st>>>(K, S->nflags.p, S->n2e.p, S->n2e_deg.p, S->nec.p, S->plan.nec_w,
        S->en0.p, S->en1.p, S->en2.p, S->surface.p, S->emass.p, S->ecbu.p, S->shape.p, S->ssh.p,
        S->drag_ui.p, S->drag_ui_young.p, S->conc.p, S->conc_young.p, S->wind.p, S->lat.p,
        S->VT[S->cur], S->VTM.p, S->node_mass.p, S->rlmass.p, S->cbu.p, S->fcor.p, S->grad_ssh.p, S->tau_a.p,
        S->ow_list.p, S->ow_count.p);
    S->n_launch += 2;
    NSX_CUDA(cudaGetLastError());
}

static void launch_tiles(nsx_solver* S, SubArgs const& A, int tile_base, int ntiles, cudaStream_t st, int max_ctas)
{
    if (ntiles <= 0) return;
    SubArgs a = A;
    a.tile_base = tile_base;
    a.n_tiles = ntiles;
    // persistent CTAs, one per SM (two shared-memory stages each); CTA b takes tiles b, b+grid, ...
    int const grid = std::max(1, std::min(ntiles, max_ctas * SUB_CTAS_PER_SM));
    size_t const smem = 128 + (size_t)SUB_STAGES * A.L.total;
    if (S->K.dynamics_type == NSX_DYN_BBM) k_subcycle<1><<<grid, SUB_TPB, smem, st>>>(S->K, a);
    else k_subcycle<0><<<grid, SUB_TPB, smem, st>>>(S->K, a);
    S->n_launch++;
}

// SMs for the boundary launch when the interior runs the tile kernel next to it.  Both are persistent kernels that
// take tiles round-robin: with B SMs the boundary chain lasts ceil(nb/B) boundary tile-times plus the NVLink round
// trip (push, flag, wait), the interior ceil((nt-nb)/(SMs-B)) tile-times.  Measured on B200 (3 km mesh on 2 GPUs,
// nb ~ 50 of 2412 tiles, sweep of B in profiles/r1_boundary_sm_sweep.txt): a boundary tile costs ~1.65 interior
// tile-times (6.2 us against 3.8 us: remote stores, larger halos, no neighbours sharing L2 lines) and the round trip
// ~4; the model below keeps a margin on the boundary side, whose penalty is the steeper one (B=4: 90 us, B=6..8:
// 70 us, B=24: 76 us per sub-cycle).  Pick the B minimising the longer of the two.
static int balanced_boundary_sms(int nb, int nt, int sms)
{
    double const boundary_tile = 1.8, latency_tiles = 5.0;
    int best = 1;
    double best_t = 1e300;
    for (int B = 1; B < sms && B <= nb; ++B) {
        double const tb = boundary_tile * std::ceil((double)nb / B) + latency_tiles;
        double const ti = std::ceil((double)(nt - nb) / (sms - B));
        double const t = std::max(tb, ti);
        if (t < best_t - 1e-9) { best_t = t; best = B; }
    }
    return best;
}

static DirectArgs direct_args(nsx_solver* S, SubArgs const& A, bool mixed)
{
    DirectArgs D{};
    D.en0 = S->en0.p; D.en1 = S->en1.p; D.en2 = S->en2.p;
    D.shape = S->shape.p; D.ec = S->ec_e.p;
    D.s0 = A.s0o; D.s1 = A.s1o; D.s2 = A.s2o; D.dm = A.dmo;
    D.s0i = A.s0i; D.s1i = A.s1i; D.s2i = A.s2i; D.di = A.di;
    D.contrib = S->contrib.p; D.elem_nowrite = mixed ? S->elem_nowrite.p : nullptr;
    D.nflags = S->nflags.p; D.n2e = S->n2e.p; D.n2e_deg = S->n2e_deg.p;
    D.grad_ssh = S->grad_ssh.p; D.node_mass = S->node_mass.p; D.rlmass = S->rlmass.p; D.cbu = S->cbu.p; D.fcor = S->fcor.p;
    D.tau_a = S->tau_a.p; D.tau_wi = A.tau_wi; D.ocean = S->ocean.p; D.VTM = S->VTM.p;
    D.disp = S->disp.p;
    return D;
}

// mailbox exchange arguments of exchange `ex` (1-based inside the model step); off for in-process groups and single ranks
static MbExchange mb_exchange(nsx_solver* S, bool on, int ex)
{
    MbExchange X{};
    X.on = on ? 1 : 0;
    X.ex = ex;
    X.mb = (MbEntry*)S->mailbox; X.n_mb = S->n_mb;
    int slot = 0;
    for (auto& p : S->peers) {
        if (p.h_send_idx.empty()) continue;
        if (slot >= MB_MAX_PEERS) throw std::runtime_error("mailbox exchange: too many neighbour ranks");
        X.peer_mb[slot] = (MbEntry*)p.peer_mb; X.peer_nmb[slot] = p.peer_nmb;
        ++slot;
    }
    X.push_ptr = S->push_ptr.p; X.push_ent = S->push_ent_mb.p;
    X.epoch_ctr = S->d_epoch.p; X.err = S->halo_err.p;
    return X;
}
static void ghost_import(nsx_solver* S, MbExchange const& X, double dt)
{
    int const nghost = S->nn - S->ndof;
    if (!X.on || nghost <= 0) return;
    k_ghost_import<<<nblk(nghost), TPB, 0, S->stream>>>(X, S->nn, S->ndof, dt, S->VT[S->cur], S->disp.p);
    S->n_launch++;
}

// one sub-cycle including its ghost exchange; flips the VT and sigma parities.
// overlap: boundary tiles first, then the halo kernel on the main stream while the interior tiles run on stream2.
static void phase_substep(nsx_solver* S, int s, bool exchange_sync, bool overlap)
{
    KParams const& K = S->K;
    // the direct path updates sigma/damage in place (no element is read by another thread): smaller L2 footprint
    bool const inplace = S->direct && S->peers.empty();
    int const so = S->scur, sn = inplace ? S->scur : (S->scur ^ 1);
    bool const bbm = (K.dynamics_type == NSX_DYN_BBM);
    int const dn = inplace ? S->dcur : (S->dcur ^ 1);
    SubArgs A{};
    A.tiles = S->tiles.p; A.tile_order = S->tile_order.p; A.tile_base = 0;
    A.halo_nodes = S->halo_nodes.p; A.halo_elems = S->halo_elems.p; A.slot_conn = S->slot_conn.p;
    A.slot_shape = S->slot_shape.p; A.slot_ec = S->slot_ec.p; A.nslots = S->plan.nslots; A.inc = S->inc.p;
    A.s0i = S->sig[so][0].p; A.s1i = S->sig[so][1].p; A.s2i = S->sig[so][2].p;
    A.s0o = S->sig[sn][0].p; A.s1o = S->sig[sn][1].p; A.s2o = S->sig[sn][2].p;
    A.di = bbm ? S->dmg[S->dcur].p : nullptr; A.dmo = bbm ? S->dmg[dn].p : nullptr;
    A.nflags = S->nflags.p; A.grad_ssh = S->grad_ssh.p; A.node_mass = S->node_mass.p; A.rlmass = S->rlmass.p;
    A.cbu = S->cbu.p; A.fcor = S->fcor.p; A.tau_a = S->tau_a.p; A.tau_wi = S->have_tau_wi ? S->tau_wi.p : nullptr;
    A.ocean = S->ocean.p; A.VTM = S->VTM.p; A.VTc = S->VT[S->cur]; A.VTn = S->VT[S->cur ^ 1];
    A.disp = S->disp.p;
    A.halo_err = S->halo_err.p;
    A.move_mesh = (K.dynamics_type != NSX_DYN_MEVP);
    A.lag_ghost_move = (A.move_mesh && s > 0);
    int npl = 0;
    for (int p = 0; p < NP_COUNT; ++p) {
        bool use = p < NP_DSU;
        if (p == NP_DSU || p == NP_DSV) use = A.move_mesh;
        if (p == NP_VMU || p == NP_VMV) use = (K.dynamics_type == NSX_DYN_MEVP);
        if (p == NP_TWU || p == NP_TWV) use = S->have_tau_wi;
        A.np[p] = use ? npl++ : 0;
    }
    A.L = sub_layout(S->plan, npl);
    if (128 + (size_t)SUB_STAGES * A.L.total > 227 * 1024) throw std::runtime_error("sub-cycle kernel: tile working set exceeds shared memory");
    S->sub_smem = (size_t)A.L.total;
    int const nt = S->plan.ntiles, nb = S->n_boundary_tiles;
    // direct path: one thread per element, then one per node.  `mixed`: the boundary tiles are handled by the tile
    // kernel of the boundary launch; here their elements are evaluated but not written and their nodes are skipped.
    auto launch_direct = [&](cudaStream_t st, bool mixed) {
        DirectArgs D = direct_args(S, A, mixed);
        int const ge = (S->ne + DIRECT_TPB - 1) / DIRECT_TPB, gn = (S->nn + DIRECT_TPB - 1) / DIRECT_TPB;
        if (bbm) k_element_direct<1><<<ge, DIRECT_TPB, 0, st>>>(K, D, A.VTc);
        else k_element_direct<0><<<ge, DIRECT_TPB, 0, st>>>(K, D, A.VTc);
        int const skip = mixed ? (NF_BTILE | NF_GHOST) : 0;
        k_node_direct<<<gn, DIRECT_TPB, 0, st>>>(K, D, A.move_mesh, mixed ? 0 : A.lag_ghost_move, skip, A.VTc, A.VTn);
        S->n_launch += 2;
    };
    // one process per GPU: mailbox exchange.  The kernels that compute the sent nodes store them straight into the holders'
    // mailboxes (boundary tiles run first); the NEXT sub-cycle's kernels read ghost velocities from this rank's mailbox
    // (polling the tag, which covers a neighbour that lags) -- no exchange kernel, no flags, no second launch chain.  After
    // the loop k_ghost_import copies the last exchange into the velocity buffer.  NsxCreateOptions.overlap = 2 selects the
    // round-1 fused boundary launch with epoch flags instead (kept for comparison).
    bool const mailbox = exchange_sync && S->halo_ready && S->opt.overlap == 1;
    if (mailbox) {
        A.X = mb_exchange(S, true, s + 1);
        if (S->direct) {
            DirectArgs D = direct_args(S, A, false);
            D.X = A.X;
            int const ge = (S->ne + DIRECT_TPB - 1) / DIRECT_TPB, gn = (S->nn + DIRECT_TPB - 1) / DIRECT_TPB;
            if (bbm) k_element_direct<1><<<ge, DIRECT_TPB, 0, S->stream>>>(K, D, A.VTc);
            else k_element_direct<0><<<ge, DIRECT_TPB, 0, S->stream>>>(K, D, A.VTc);
            k_node_direct<<<gn, DIRECT_TPB, 0, S->stream>>>(K, D, A.move_mesh, A.lag_ghost_move, 0, A.VTc, A.VTn);
            S->n_launch += 2;
        } else {
            launch_tiles(S, A, 0, nt, S->stream, S->sm_count);
        }
        S->cur ^= 1;
        S->scur = sn;
        if (bbm) S->dcur = dn;
        NSX_CUDA(cudaGetLastError());
        return;
    }
    bool const fused = overlap && exchange_sync && nb > 0 && nb < nt && S->halo_ready;
    if (fused) {
        // multi-GPU sub-cycle: [boundary tiles + NVLink push + epoch signal + wait] as ONE kernel on a few SMs of the
        // main stream, the interior (tile kernel or direct kernels) concurrently on stream2 on the remaining SMs
        NSX_CUDA(cudaEventRecord(S->ev_fork, S->stream));
        NSX_CUDA(cudaStreamWaitEvent(S->stream2, S->ev_fork, 0));
        // SM split between the two co-resident persistent kernels.  Direct path (small meshes): the boundary chain
        // (tiles + NVLink round trip) is latency-bound and on the critical path of both ranks of a pair: one tile per
        // CTA.  Tile path: the split that lets both kernels finish together (balanced_boundary_sms).
        int B = S->direct ? std::min(nb, 48) : balanced_boundary_sms(nb, nt, S->sm_count);
        if (S->opt.boundary_sms > 0) B = std::max(1, std::min(S->opt.boundary_sms, std::min(nb, S->sm_count - 1)));
        SubArgs Ab = A;
        Ab.fuse_halo = 1;
        Ab.H = halo_args(S, S->cur ^ 1, true);
        Ab.push_ptr = S->push_ptr.p; Ab.push_ent = S->push_ent.p;
        Ab.my_flags = S->flags; Ab.epoch_ctr = S->d_epoch.p; Ab.done_ctr = S->d_done.p; Ab.halo_err = S->halo_err.p;
        launch_tiles(S, Ab, 0, nb, S->stream, B);
        if (S->direct) launch_direct(S->stream2, true);
        else launch_tiles(S, A, nb, nt - nb, S->stream2, S->sm_count - B);
        NSX_CUDA(cudaEventRecord(S->ev_join, S->stream2));
        S->cur ^= 1;
        NSX_CUDA(cudaStreamWaitEvent(S->stream, S->ev_join, 0));
    } else if (S->direct) {
        launch_direct(S->stream, false);
        S->cur ^= 1;
        if (exchange_sync) halo_exchange(S, true);
    } else {
        launch_tiles(S, A, 0, nt, S->stream, S->sm_count);
        S->cur ^= 1;
        if (exchange_sync) halo_exchange(S, true);
    }
    S->scur = sn;
    if (bbm) S->dcur = dn;              // EVP / mEVP never touch damage (FE.cpp:10649-10699)
    NSX_CUDA(cudaGetLastError());
}

// one process per GPU on the tile / direct paths with the mailbox exchange (ghosts are imported and moved every sub-cycle)
static bool mailbox_mode(nsx_solver const* S)
{
    return !S->resident && !S->halo_local && !S->peers.empty() && S->halo_ready && S->opt.overlap == 1;
}

// mesh moves that follow the loop: mEVP's single move (FE.cpp:10559-10573) or the ghosts' lagged last move
static void phase_post_move(nsx_solver* S, int nrun)
{
    cudaStream_t st = S->stream;
    KParams const& K = S->K;
    if (K.dynamics_type == NSX_DYN_MEVP) {
        k_move_mesh<<<nblk(S->nn), TPB, 0, st>>>(S->nn, 0, S->nn, K.dtime_step, S->VT[S->cur], S->disp.p);
        S->n_launch++;
    } else if (S->nn > S->ndof && nrun > 0) {
        k_move_mesh<<<nblk(S->nn - S->ndof), TPB, 0, st>>>(S->nn, S->ndof, S->nn, K.dte, S->VT[S->cur], S->disp.p);
        S->n_launch++;
    }
    NSX_CUDA(cudaGetLastError());
}

// Both ping-pong buffers must agree on the nodes a sweep does not write.  Only OWNED entries are copied:
// ghost slots of the other buffer belong to the owners' pushes (a faster peer may already be writing them).
static void phase_ow_begin(nsx_solver* S)
{
    size_t const nn = S->nn, nd = S->ndof;
    NSX_CUDA(cudaMemcpyAsync(S->VT[S->cur ^ 1], S->VT[S->cur], nd * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
    NSX_CUDA(cudaMemcpyAsync(S->VT[S->cur ^ 1] + nn, S->VT[S->cur] + nn, nd * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
}
static void phase_ow_sweep(nsx_solver* S)
{
    int const grid = std::min(nblk(S->ndof), S->sm_count * 4);
    k_ow_sweep<<<grid, TPB, 0, S->stream>>>(S->nn, S->ow_list.p, S->ow_count.p, S->n2n.p, S->n2n_deg.p,
                                             S->VT[S->cur], S->VT[S->cur ^ 1]);
    S->n_launch++;
    S->cur ^= 1;
    NSX_CUDA(cudaGetLastError());
}
static void phase_tauw(nsx_solver* S)
{
    // the resident launch reads the exchange epoch at its start; the counter advances here, after it has finished
    k_tauw_owmove<<<nblk(S->nn), TPB, 0, S->stream>>>(S->K, S->nflags.p, S->node_mass.p, S->VT[S->cur], S->VTM.p,
                                                      S->ocean.p, S->tau_w.p, S->UM.p, S->UT.p, S->resident ? nullptr : S->disp.p,
                                                      S->epoch_bump ? S->d_epoch.p : nullptr, (unsigned long long)S->epoch_bump);
    S->epoch_bump = 0;
    S->n_launch++;
    NSX_CUDA(cudaGetLastError());
}

static int substeps_to_run(nsx_solver const* S)
{
    int const steps = S->P.substeps;
    return (S->P.stop_after_substeps > 0) ? std::min(steps, S->P.stop_after_substeps) : steps;
}

// timing events: inside a stream capture they must be recorded as external event nodes
static void record(nsx_solver* S, int i)
{
    if (S->capturing) NSX_CUDA(cudaEventRecordWithFlags(S->ev[i], S->stream, cudaEventRecordExternal));
    else NSX_CUDA(cudaEventRecord(S->ev[i], S->stream));
}

// The whole sub-cycle loop and the open-water smoother of one rank in ONE cooperative launch (k_resident).
static void launch_resident(nsx_solver* S, int nrun, bool cooperative)
{
    KParams const& K = S->K;
    bool const bbm = (K.dynamics_type == NSX_DYN_BBM);
    int const nsweeps = S->P.skip_ow_smoother ? 0 : 50;            // hard-coded 50 sweeps, FE.cpp:10580
    if (!S->peers.empty() && !S->halo_ready) throw std::runtime_error("explicit solve before nsx_halo_finalize");
    ResidentArgs A{};
    A.tiles = S->tiles.p; A.rtiles = S->res_tiles.p;
    A.halo_nodes = S->halo_nodes.p; A.halo_slot = S->halo_slot.p; A.halo_move = S->halo_move.p; A.halo_elems = S->halo_elems.p;
    A.slot_conn = S->slot_conn.p;
    A.slot_shape = S->slot_shape.p; A.slot_ec = S->slot_ec.p; A.nslots = S->plan.nslots; A.inc = S->inc.p;
    A.n2n_loc = S->res_n2n.p; A.n2n_deg = S->res_n2n_deg.p;
    A.s0 = S->sig[S->scur][0].p; A.s1 = S->sig[S->scur][1].p; A.s2 = S->sig[S->scur][2].p;
    A.dm = bbm ? S->dmg[S->dcur].p : nullptr;
    A.nflags = S->nflags.p; A.grad_ssh = S->grad_ssh.p; A.node_mass = S->node_mass.p; A.rlmass = S->rlmass.p;
    A.cbu = S->cbu.p; A.fcor = S->fcor.p; A.tau_a = S->tau_a.p; A.tau_wi = S->have_tau_wi ? S->tau_wi.p : nullptr;
    A.ocean = S->ocean.p; A.VTM = S->VTM.p;
    A.VT0 = S->VT[0]; A.VT1 = S->VT[1]; A.cur = S->cur; A.UM = S->UM.p; A.UT = S->UT.p;
    A.move_mesh = (K.dynamics_type != NSX_DYN_MEVP); A.nsub = nrun; A.nsweeps = nsweeps;
    A.ow_count = S->ow_count.p;
    A.mb = (MbEntry*)S->mailbox; A.n_mb = S->n_mb;
    A.push_ptr = S->push_ptr.p; A.push_ent = S->push_ent_mb.p;
    A.epoch_ctr = S->d_epoch.p; A.err = S->halo_err.p;
    S->res_time.zero(S->stream);
    A.tstamp = S->res_time.p;
    A.MS = S->plan.msp; A.MLN = S->plan.max_local_nodes + 2;
    int slot = 0;
    for (auto& p : S->peers) {
        if (p.h_send_idx.empty()) continue;
        A.P.send_mb[slot] = (MbEntry*)p.peer_mb; A.P.send_nmb[slot] = p.peer_nmb;
        ++slot;
    }
    A.P.n_send = slot;
    A.has_peers = S->peers.empty() ? 0 : 1;
    size_t const smem = (size_t)((bbm ? 16 : 12) * A.MS + 2 * A.MLN) * sizeof(double);
    // cooperative launch: the driver guarantees that all CTAs (one per SM) are co-resident, which the flag protocol needs
    void* args[2] = {(void*)&K, (void*)&A};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(S->plan.ntiles); cfg.blockDim = dim3(RES_TPB); cfg.dynamicSmemBytes = smem; cfg.stream = S->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    // ranks sharing one GPU (in-process groups of the tests) launch plainly: their kernels must overlap each other, the
    // host has checked that together they fit the device, and every spin is bounded
    cfg.attrs = at; cfg.numAttrs = cooperative ? 1 : 0;
    if (bbm) NSX_CUDA(cudaLaunchKernelExC(&cfg, (const void*)k_resident<1>, args));
    else NSX_CUDA(cudaLaunchKernelExC(&cfg, (const void*)k_resident<0>, args));
    S->n_launch++;
    S->cur = (S->cur + nrun + nsweeps) & 1;          // the kernel leaves the result in that buffer whether or not it swept
    S->epoch_bump = nrun + nsweeps;                  // mailbox tags never repeat: the epoch advances after every launch
}

static void solve_group(int n, nsx_solver** W)
{
    // lock-step over ranks; for n == 1 this is the plain single-rank sequence.  Ranks in a group are
    // serialised through events at exchange points; with one process per GPU (n == 1, peers remote) the
    // exchange waits on NVLink flags instead.
    for (int r = 0; r < n; ++r) {
        nsx_solver* S = W[r];
        if (!S->have_params) throw std::runtime_error("nsx_explicit_solve before nsx_set_params");
        NSX_CUDA(cudaSetDevice(S->device));
        S->n_launch = 0;
        record(S, 0);
        phase_prep(S);
        record(S, 1);
    }
    int const nrun = substeps_to_run(W[0]);
    bool const remote = (n == 1) && !W[0]->halo_local && !W[0]->peers.empty();
    bool const overlap_on = W[0]->opt.overlap != 0;
    bool const ow_skip = W[0]->opt.ow_skip != 0;
    auto group_exchange = [&]() {
        // in-process group: all pushes, then a cross-stream join so nobody reads ghosts too early
        for (int r = 0; r < n; ++r) { NSX_CUDA(cudaSetDevice(W[r]->device)); halo_exchange(W[r], false); NSX_CUDA(cudaEventRecord(W[r]->ev[5], W[r]->stream)); }
        for (int r = 0; r < n; ++r)
            for (int q = 0; q < n; ++q)
                if (q != r) NSX_CUDA(cudaStreamWaitEvent(W[r]->stream, W[q]->ev[5], 0));
    };
    bool const resident = W[0]->resident;
    if (resident) {
        // state-resident path: the sub-cycle loop AND the smoother of every rank are one persistent launch each.  Ranks
        // of an in-process group run concurrently on their own streams and synchronise through the same flags as ranks
        // on different GPUs, so together they must fit the device (NsxCreateOptions.max_sms).
        int tiles_on_dev = 0;
        for (int r = 0; r < n; ++r) {
            if (!W[r]->resident) throw std::runtime_error("group solve: every rank must use the same path");
            if (W[r]->device == W[0]->device) tiles_on_dev += W[r]->plan.ntiles;
        }
        if (n > 1 && tiles_on_dev > RES_CTAS * W[0]->sm_count)
            throw std::runtime_error("group solve: the resident launches of the ranks sharing this GPU need " + std::to_string(tiles_on_dev) +
                                     " SMs; create the handles with NsxCreateOptions.max_sms = SMs / ranks");
        for (int r = 0; r < n; ++r) { NSX_CUDA(cudaSetDevice(W[r]->device)); launch_resident(W[r], nrun, n == 1); }
    } else {
        for (int s = 0; s < nrun; ++s) {
            for (int r = 0; r < n; ++r) { NSX_CUDA(cudaSetDevice(W[r]->device)); phase_substep(W[r], s, remote, remote && overlap_on); }
            if (n > 1) group_exchange();
        }
    }
    for (int r = 0; r < n; ++r) {
        NSX_CUDA(cudaSetDevice(W[r]->device));
        // mailbox exchange: the ghosts of the last sub-cycle go from the mailbox into the velocity buffer
        if (n == 1 && mailbox_mode(W[r])) ghost_import(W[r], mb_exchange(W[r], true, nrun), 0.);
        record(W[r], 2);
        if (!resident) phase_post_move(W[r], nrun);
    }
    if (!W[0]->P.skip_ow_smoother && !resident) {
        for (int r = 0; r < n; ++r) { NSX_CUDA(cudaSetDevice(W[r]->device)); phase_ow_begin(W[r]); }
        if (n == 1 && W[0]->peers.empty()) {
            // no neighbours: all 50 sweeps (hard-coded count, FE.cpp:10580) in one launch with a grid barrier
            nsx_solver* S = W[0];
            S->d_done.zero(S->stream);
            int const grid = std::max(1, std::min(nblk(S->ndof), S->sm_count));
            k_ow_smooth_all<<<grid, TPB, 0, S->stream>>>(S->nn, 50, S->ow_list.p, S->ow_count.p, S->n2n.p, S->n2n_deg.p,
                                                         S->VT[S->cur], S->VT[S->cur ^ 1], S->d_done.p, S->halo_err.p);
            S->n_launch++;
            NSX_CUDA(cudaGetLastError());
        } else {
            for (int nit = 0; nit < 50; ++nit) {       // hard-coded 50 sweeps, FE.cpp:10580
                if (n == 1 && remote && mailbox_mode(W[0])) {
                    // one process per GPU: sweep, push of every sent node, import of every ghost (mailbox exchange)
                    nsx_solver* S = W[0];
                    MbExchange X = mb_exchange(S, true, nrun + nit + 1);
                    int const sweep_blocks = std::max(1, std::min(nblk(S->ndof), S->sm_count * 4));
                    k_ow_sweep_mb<<<sweep_blocks + nblk(S->n_send_total), TPB, 0, S->stream>>>(X, nit == 0 ? 1 : 0, sweep_blocks, S->nn, S->ndof,
                        S->ow_list.p, S->ow_count.p, S->n2n.p, S->n2n_deg.p, S->nflags.p, S->node_mass.p, S->VT[S->cur], S->VT[S->cur ^ 1],
                        S->n_send_total, S->d_send_src.p, S->d_send_slot.p);
                    S->n_launch++;
                    S->cur ^= 1;
                    if (nit == 49) ghost_import(S, X, 0.);       // the last sweep's ghosts go into the velocity buffer
                    NSX_CUDA(cudaGetLastError());
                    continue;
                }
                if (n == 1 && remote) {
                    // round-1 protocol (NsxCreateOptions.overlap = 2): sweep + ghost exchange with epoch flags in one launch
                    nsx_solver* S = W[0];
                    HaloArgs a = halo_args(S, S->cur ^ 1, true);
                    int const grid = std::max(1, std::min(nblk(S->ndof), S->sm_count));
                    int const mode = !ow_skip ? 2 : (nit == 0) ? 0 : (nit == 49) ? 2 : 1;
                    k_ow_sweep_exchange<<<grid, TPB, 0, S->stream>>>(a, mode, S->nn, S->ow_list.p, S->ow_count.p, S->n2n.p, S->n2n_deg.p,
                        S->VT[S->cur], S->VT[S->cur ^ 1], S->d_send_src.p, S->d_send_dst.p, S->push_ptr.p, S->push_ent.p,
                        S->ow_pair.p, S->ow_pair.p + 32, S->flags, S->d_epoch.p, S->d_done.p, 40000000LL, S->halo_err.p);
                    S->n_launch++;
                    S->cur ^= 1;
                    NSX_CUDA(cudaGetLastError());
                    continue;
                }
                for (int r = 0; r < n; ++r) { NSX_CUDA(cudaSetDevice(W[r]->device)); phase_ow_sweep(W[r]); }
                if (n > 1) group_exchange();
            }
        }
    }
    if (n == 1 && mailbox_mode(W[0])) W[0]->epoch_bump = nrun + (W[0]->P.skip_ow_smoother ? 0 : 50);   // tags never repeat
    for (int r = 0; r < n; ++r) {
        nsx_solver* S = W[r];
        NSX_CUDA(cudaSetDevice(S->device));
        phase_tauw(S);
        record(S, 3);
        S->timing.n_launches = S->n_launch;
        S->timing.n_substeps = nrun;
        S->timing_valid = true;
    }
}

// explicitSolve() of one rank.  The whole launch sequence (2 prep kernels, `substeps` x {tile kernel[, halo]},
// 50 smoother sweeps, tau_w) is captured once into a CUDA graph per entry parity of the ping-pong buffers and
// replayed afterwards: at 2e5 elements a sub-cycle is a few microseconds of device work, less than the host
// cost of launching its kernels one by one.  NsxCreateOptions.use_graph = 0 disables this (debugging).
extern "C" int nsx_explicit_solve(nsx_handle S)
{
    NSX_API_BEGIN(S)
    nsx_solver* W[1] = {S};
    if (!S->opt.use_graph) {
        solve_group(1, W);
    } else {
        if (!S->graph_valid) {
            for (auto& g : S->graph_exec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
            S->graph_valid = true;
        }
        int const slot = S->cur + 2 * S->scur + 4 * S->dcur;
        int const cur_in = S->cur, scur_in = S->scur, dcur_in = S->dcur;
        if (!S->graph_exec[slot]) {
            if (!S->have_params) throw std::runtime_error("nsx_explicit_solve before nsx_set_params");
            cudaGraph_t graph = nullptr;
            NSX_CUDA(cudaStreamBeginCapture(S->stream, cudaStreamCaptureModeThreadLocal));
            S->capturing = true;
            try {
                solve_group(1, W);
            } catch (...) {
                S->capturing = false;
                cudaStreamEndCapture(S->stream, &graph);
                if (graph) cudaGraphDestroy(graph);
                S->cur = cur_in; S->scur = scur_in; S->dcur = dcur_in;
                throw;
            }
            S->capturing = false;
            NSX_CUDA(cudaStreamEndCapture(S->stream, &graph));
            cudaError_t e = cudaGraphInstantiate(&S->graph_exec[slot], graph, 0);
            cudaGraphDestroy(graph);
            NSX_CUDA(e);
            S->graph_cur_out[slot] = S->cur;
            S->graph_scur_out[slot] = S->scur;
            S->graph_dcur_out[slot] = S->dcur;
            S->graph_launches[slot] = S->timing.n_launches;
            S->graph_nsub[slot] = S->timing.n_substeps;
            S->cur = cur_in; S->scur = scur_in; S->dcur = dcur_in;
        }
        NSX_CUDA(cudaGraphLaunch(S->graph_exec[slot], S->stream));
        S->cur = S->graph_cur_out[slot];
        S->scur = S->graph_scur_out[slot];
        S->dcur = S->graph_dcur_out[slot];
        S->timing.n_launches = S->graph_launches[slot];
        S->timing.n_substeps = S->graph_nsub[slot];
        S->timing_valid = true;
    }
    NSX_API_END(S)
}

extern "C" int nsx_group_explicit_solve(int n, nsx_handle* hs)
{
    if (n <= 0 || !hs || !hs[0]) return 1;
    try {
        for (int r = 0; r < n; ++r) hs[r]->halo_local = true;
        solve_group(n, hs);
    } catch (std::exception const& e) { hs[0]->err = e.what(); return 2; }
    return 0;
}

extern "C" int nsx_update(nsx_handle S)
{
    NSX_API_BEGIN(S)
    if (!S->have_params) throw std::runtime_error("nsx_update before nsx_set_params");
    NSX_CUDA(cudaEventRecord(S->ev_upd, S->stream));      // own start event: ev[3] stays the end of explicitSolve()
    int const c = S->scur;
    k_update<<<nblk(S->ne), TPB, 0, S->stream>>>(S->K, S->nflags.p, S->en0.p, S->en1.p, S->en2.p, S->x.p, S->y.p, S->UM.p,
        S->surface.p, S->conc.p, S->thick.p, S->snow.p, S->thick_myi.p, S->conc_myi.p, S->ridge_ratio.p,
        S->conc_young.p, S->h_young.p, S->hs_young.p, S->sig[c][0].p, S->sig[c][1].p, S->sig[c][2].p, S->del_ci_ridge_myi.p);
    NSX_CUDA(cudaGetLastError());
    NSX_CUDA(cudaEventRecord(S->ev[4], S->stream));
    S->update_timed = true;
    NSX_API_END(S)
}

extern "C" int nsx_get_timing(nsx_handle S, NsxTiming* out)
{
    NSX_API_BEGIN(S)
    if (!out) throw std::invalid_argument("nsx_get_timing: NULL");
    NSX_CUDA(cudaStreamSynchronize(S->stream));
    if (S->timing_valid) {
        NSX_CUDA(cudaEventElapsedTime(&S->timing.prep_ms, S->ev[0], S->ev[1]));
        NSX_CUDA(cudaEventElapsedTime(&S->timing.subcycle_ms, S->ev[1], S->ev[2]));
        NSX_CUDA(cudaEventElapsedTime(&S->timing.ow_smoother_ms, S->ev[2], S->ev[3]));
        if (S->resident) {
            // one launch holds the sub-cycle loop AND the smoother: the CUDA-event time of the launch is split by the
            // %globaltimer stamps the kernel took at its start, after its last sub-cycle and at its end (max over tiles)
            NSX_CUDA(cudaMemcpy(S->h_time, S->res_time.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            double const all = (double)(S->h_time[2] - S->h_time[0]), loop = (double)(S->h_time[1] - S->h_time[0]);
            if (S->h_time[0] && all > 0. && loop > 0. && loop <= all) {
                float const launch_ms = S->timing.subcycle_ms;
                S->timing.subcycle_ms = (float)(launch_ms * loop / all);
                S->timing.ow_smoother_ms += launch_ms - S->timing.subcycle_ms;
            }
        }
    }
    if (S->update_timed) NSX_CUDA(cudaEventElapsedTime(&S->timing.update_ms, S->ev_upd, S->ev[4]));
    *out = S->timing;
    NSX_API_END(S)
}

extern "C" int nsx_check(nsx_handle S, NsxCheck* out)
{
    NSX_API_BEGIN(S)
    if (!out) throw std::invalid_argument("nsx_check: NULL");
    S->check_i.zero(S->stream); S->check_d.zero(S->stream);
    int const n = std::max(S->nn, S->ne);
    int const c = S->scur;
    k_check<<<nblk(n), TPB, 0, S->stream>>>(S->nn, S->ndof, S->ne, S->VT[S->cur], S->sig[c][0].p, S->sig[c][1].p, S->sig[c][2].p,
                                            S->dmg[S->dcur].p, S->conc.p, S->thick.p, S->check_i.p, (unsigned long long*)S->check_d.p);
    NSX_CUDA(cudaGetLastError());
    int hi[4]; double hd;
    NSX_CUDA(cudaMemcpyAsync(hi, S->check_i.p, sizeof(hi), cudaMemcpyDeviceToHost, S->stream));
    NSX_CUDA(cudaMemcpyAsync(&hd, S->check_d.p, sizeof(hd), cudaMemcpyDeviceToHost, S->stream));
    NSX_CUDA(cudaStreamSynchronize(S->stream));
    out->n_nan = hi[0]; out->n_speed = hi[1]; out->n_range = hi[2]; out->pad_ = 0; out->max_speed = hd;
    NSX_API_END(S)
}

// ---------------------------------------------------------------------------------------------------
// regrid check, ice diagnostics and forcing interpolation on the resident state
// ---------------------------------------------------------------------------------------------------
static double key_to_double(unsigned long long k)
{
    unsigned long long const b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    double v;
    std::memcpy(&v, &b, sizeof(v));
    return v;
}

extern "C" int nsx_check_regridding(nsx_handle S, double regrid_angle, NsxRegrid* out)
{
    NSX_API_BEGIN(S)
    if (!out) throw std::invalid_argument("nsx_check_regridding: NULL");
    if (!S->regrid_keys.p) S->regrid_keys.alloc(4);
    unsigned long long const init[4] = {~0ULL, ~0ULL, 0ULL, 0ULL};
    NSX_CUDA(cudaMemcpyAsync(S->regrid_keys.p, init, sizeof(init), cudaMemcpyHostToDevice, S->stream));
    int const grid = std::max(1, std::min(nblk(S->ne), 8 * S->sm_count));
    k_regrid_check<<<grid, TPB, 0, S->stream>>>(S->ne, S->nn, S->en0.p, S->en1.p, S->en2.p, S->x.p, S->y.p, S->UM.p,
                                                S->regrid_keys.p);
    NSX_CUDA(cudaGetLastError());
    unsigned long long k[4];
    NSX_CUDA(cudaMemcpyAsync(k, S->regrid_keys.p, sizeof(k), cudaMemcpyDeviceToHost, S->stream));
    NSX_CUDA(cudaStreamSynchronize(S->stream));
    out->min_angle = key_to_double(k[0]);
    out->min_jacobian = key_to_double(k[1]);
    out->max_jacobian = key_to_double(k[2]);
    out->flip = (out->min_jacobian <= 0.) && (out->max_jacobian >= 0.);
    out->regrid = (out->min_angle < regrid_angle) || out->flip;
    NSX_API_END(S)
}

extern "C" int nsx_update_ice_diagnostics(nsx_handle S)
{
    NSX_API_BEGIN(S)
    if (!S->have_params) throw std::runtime_error("nsx_update_ice_diagnostics before nsx_set_params");
    if (!S->diag.p) S->diag.alloc(6 * (size_t)S->ne);
    int const c = S->scur;
    k_ice_diagnostics<<<nblk(S->ne), TPB, 0, S->stream>>>(S->ne, S->nn, S->K.young_ice, S->en0.p, S->en1.p, S->en2.p,
        S->x.p, S->y.p, S->UM.p, S->VT[S->cur], S->conc.p, S->thick.p, S->snow.p, S->conc_young.p, S->h_young.p,
        S->hs_young.p, S->sig[c][0].p, S->sig[c][1].p, S->sig[c][2].p, S->diag.p);
    NSX_CUDA(cudaGetLastError());
    NSX_API_END(S)
}

static int forcing_planes(int var) { return var == NSX_FORCING_SSH ? 1 : 2; }

extern "C" int nsx_forcing_load(nsx_handle S, int var, int slot, const double* data)
{
    NSX_API_BEGIN(S)
    if (var < 0 || var > 2 || slot < 0 || slot > 1 || !data) throw std::invalid_argument("nsx_forcing_load: bad argument");
    int const planes = forcing_planes(var);
    auto& buf = S->forcing[var][slot];
    if (!buf.p) buf.alloc((size_t)planes * S->nn);
    if (S->arena.n < (size_t)planes * S->nn) { NSX_CUDA(cudaStreamSynchronize(S->stream)); S->arena.alloc((size_t)planes * S->nn); }
    NSX_CUDA(cudaMemcpyAsync(S->arena.p, data, (size_t)planes * S->nn * sizeof(double), cudaMemcpyHostToDevice, S->stream));
    k_permute_in<<<nblk(S->nn), TPB, 0, S->stream>>>(S->nn, planes, S->node_perm.p, S->arena.p, buf.p);
    NSX_CUDA(cudaGetLastError());
    NSX_CUDA(cudaStreamSynchronize(S->stream));             // the caller may reuse `data` right away
    S->forcing_loaded[var][slot] = true;
    NSX_API_END(S)
}

extern "C" int nsx_forcing_apply(nsx_handle S, int var, int interp_linear_time, double current_time, double ftime0,
                                 double ftime1, double factor, double bias_correction)
{
    NSX_API_BEGIN(S)
    if (var < 0 || var > 2) throw std::invalid_argument("nsx_forcing_apply: unknown variable");
    if (!S->forcing_loaded[var][0] || (interp_linear_time && !S->forcing_loaded[var][1]))
        throw std::runtime_error("nsx_forcing_apply: time slice not loaded (nsx_forcing_load)");
    double c0 = 1., c1 = 0.;
    if (interp_linear_time) {                               // externaldata.cpp:368-370
        double const fdt = std::fabs(ftime1 - ftime0);
        c0 = std::fabs(current_time - ftime1) / fdt;
        c1 = std::fabs(current_time - ftime0) / fdt;
    }
    long const n = (long)forcing_planes(var) * S->nn;
    double* const dst = var == NSX_FORCING_WIND ? S->wind.p : var == NSX_FORCING_OCEAN ? S->ocean.p : S->ssh.p;
    double const* const d0 = S->forcing[var][0].p;
    double const* const d1 = interp_linear_time ? S->forcing[var][1].p : d0;
    k_forcing_apply<<<nblk(n), TPB, 0, S->stream>>>(n, interp_linear_time, c0, c1, factor, bias_correction, d0, d1, dst);
    NSX_CUDA(cudaGetLastError());
    NSX_API_END(S)
}

// Host-only: builds the tile plan for a mesh exactly like nsx_create would (same shrink-to-fit loop) and returns
// its statistics without touching the GPU.  out[0..9] = ntiles, nodes/tile, nslots, max_local_nodes, max_slots,
// max_own_slots, max_halo_slots, max_halo_nodes, stage bytes (14 node planes), shrink attempts, tiles reading ghosts.
extern "C" int nsx_plan_info(const NsxMesh* mesh, int target_tile_nodes, int wave_ctas, int* out, int n)
{
    try {
        MeshPlan P;
        bool const no_shrink = target_tile_nodes < 0;          // resident path: exactly the requested tile size
        int target = target_tile_nodes > 0 ? target_tile_nodes : (no_shrink ? -target_tile_nodes : 208);
        int attempt = 0;
        for (;; ++attempt) {
            P = MeshPlan();
            build_mesh_plan(mesh, P, target, wave_ctas);
            if (no_shrink || sub_layout(P, 14).total <= SUB_SMEM_CAP) break;
            if (attempt > 12 || target <= 32) break;
            target = std::max(32, (int)(target * 0.88));
        }
        int nbt = 0;
        for (auto const& td : P.tiles) nbt += td.boundary;      // tiles that read ghost nodes
        int const v[11] = {P.ntiles, P.tile_nodes, P.nslots, P.max_local_nodes, P.max_slots, P.max_own_slots,
                           P.max_halo_slots, P.max_halo_nodes, sub_layout(P, 14).total, attempt, nbt};
        for (int i = 0; i < n && i < 11; ++i) out[i] = v[i];
        return 0;
    } catch (std::exception const& e) {
        g_create_err = e.what();
        return 2;
    }
}

// host only (no GPU): the state-resident plan nsx_create_ex would try for this rank on `sms` SMs.
// out[0..11] = fits, ntiles, nodes/tile, slot space, max slots per tile, max local nodes per tile, shared-memory bytes
// (BBM), export nodes, early own slots (sum), halo slots (sum), own slots (sum), limit of shared-memory bytes
extern "C" int nsx_resident_plan_info(const NsxMesh* mesh, const NsxHalo* halo, int sms, int* out, int n)
{
    try {
        if (!mesh) throw std::invalid_argument("nsx_resident_plan_info: NULL mesh");
        validate_inputs(mesh, halo);
        MeshPlan P;
        std::string why;
        bool const fits = try_resident_plan(mesh, halo, RES_CTAS * std::max(1, sms), P, why);
        if (!fits) g_create_err = why;
        long early = 0, halo_slots = 0, own = 0;
        for (size_t t = 0; t < P.tiles.size(); ++t) {
            own += P.tiles[t].n_own_slots; halo_slots += P.tiles[t].n_halo_slots;
            if (t < P.res_tiles.size()) early += P.res_tiles[t].n_early_own;
        }
        int const v[14] = {fits ? 1 : 0, P.ntiles, P.tile_nodes, P.nslots, P.max_slots, P.max_local_nodes,
                           (int)((size_t)(16 * P.msp + 2 * (P.max_local_nodes + 2)) * sizeof(double)), P.n_export,
                           (int)early, (int)halo_slots, (int)own, RES_SMEM_MAX, (int)P.p2_wavefronts, (int)P.p2_cells};
        for (int i = 0; i < n && i < 14; ++i) out[i] = v[i];
        return 0;
    } catch (std::exception const& e) {
        g_create_err = e.what();
        return 2;
    }
}

// ---------------------------------------------------------------------------------------------------
// thermo()
// ---------------------------------------------------------------------------------------------------
#include "nsx_thermo_api.cuh"

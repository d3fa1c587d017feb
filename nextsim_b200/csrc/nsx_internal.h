// nsx_internal.h -- private state of one solver handle (one rank == one GPU).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nsx.h"
#include "nsx_mesh.h"
#include "nsx_thermo.cuh"

namespace nsx {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define NSX_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            throw nsx::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e_) +   \
                                 " (" __FILE__ ":" + std::to_string(__LINE__) + ")");          \
    } while (0)

// physical constants, model/constants.hpp:56-86
constexpr double RHOI = 917., RHOW = 1025., RHOS = 330., GRAVITY = 9.80616, OMEGA = 7.292e-5, RHOA = 1.22;
constexpr double PI_ = 3.14159265358979323846;
constexpr double DAYS_IN_SEC = 86400.;

// node flag bits (also written by nsx_mesh.cpp)
enum : uint8_t { NF_DIRICHLET = 1, NF_NEUMANN = 2, NF_GHOST = 4, NF_LATNEG = 8, NF_BTILE = 16 };

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) NSX_CUDA(cudaMalloc(&p, n * sizeof(T) + 32));     // 32 B slack: TMA copies round up to 16 B granules
    }
    void zero(cudaStream_t s) { if (n) NSX_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(std::vector<T> const& h, cudaStream_t s) {
        alloc(h.size());
        if (n) NSX_CUDA(cudaMemcpyAsync(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DBuf() { release(); }
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
};

// Scalars the kernels need, derived once from NsxDynParams (passed by value to kernels).
struct KParams {
    int dynamics_type, basal_stress_type, young_ice;
    int nn, ndof, ne;
    double dte, dtime_step;
    double cos_ota, sin_ota_abs, min_m;
    // BBM
    double young, compaction_param, lambda0, exp_relax_m1, compression_factor, exp_compression;
    double compr_strength, tan_phi, sqrt_nu_rhoi;
    double D00, D01, D22;                  // M_Dunit non-zeros (FE.cpp:1491-1507)
    int    relax_int_pow;                  // exponent_relaxation_sigma-1 if a small integer, else -1
    // EVP / mEVP
    double evp_e, evp_Pstar, evp_C, evp_dmin, ralpha1, ralpha2, re2;
    double mevp_b, mevp_rb, dte_mevp;      // beta+1, 1/(beta+1), dte/(beta+1)
    // nodal solve
    double rhow_cdw, u0;
    // basal
    double k1, k2, Cb;
    // update()
    double min_c, min_h;
    int equal_ridging, myi_with_young;
};

struct PeerLink {
    int rank = -1;
    std::vector<int> h_send_idx;           // my INTERNAL node ids to send (M_extract_local_index[peer], permuted)
    std::vector<int> h_recv_idx;           // my INTERNAL ghost ids filled by that peer (M_local_ghosts_local_index[peer])
    std::vector<int> h_send_dst;           // the peer's internal ghost ids for my send list (same order)
    std::vector<int> h_send_slot;          // the peer's MAILBOX slots for my send list (resident path)
    void* peer_mb = nullptr;               // peer's mailbox (mapped), parity 0
    int peer_nmb = 0;                      // entries per parity of the peer's mailbox
    double* peer_vt[2] = {nullptr, nullptr};   // peer's VT ping-pong buffers (mapped)
    unsigned long long* peer_flags = nullptr;  // peer's flag array (mapped); I write slot [my rank]
    int peer_nn = 0;
    void* ipc_base = nullptr;              // cudaIpcOpenMemHandle result (to close)
    bool connected = false;
};

} // namespace nsx

struct nsx_solver {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;            // interior tiles while the halo of the boundary tiles is in flight
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;

    int nn = 0, ndof = 0, ne = 0, ne_local = 0;
    int rank = 0, nranks = 1;
    NsxDynParams P{};
    bool have_params = false;
    nsx::KParams K{};

    // ---- mesh (device, internal numbering) ----
    nsx::MeshPlan plan;                        // host copy (permutations are needed for the halo wiring)
    nsx::DBuf<int> node_perm, elem_perm;       // reference id -> internal id
    nsx::DBuf<double> x, y, lat;
    nsx::DBuf<uint8_t> nflags;
    nsx::DBuf<int> en0, en1, en2;
    nsx::DBuf<int> n2e, n2e_deg, nec, n2n, n2n_deg;
    // tiles
    nsx::DBuf<nsx::TileDesc> tiles;
    nsx::DBuf<int> tile_order;                 // boundary tiles first
    int n_boundary_tiles = 0;
    nsx::DBuf<int> halo_nodes, halo_elems, slot_elem;
    nsx::DBuf<unsigned long long> slot_conn;
    nsx::DBuf<uint16_t> inc;
    nsx::DBuf<double> slot_shape, slot_ec;     // slot space: 4 shape planes, 6 rheology-constant planes (2 for EVP / mEVP)
    size_t sub_smem = 0;

    // ---- fields (device) ----
    // halo window: VT ping-pong buffers + flags live in ONE allocation so it can be IPC-exported
    void* window = nullptr;
    size_t window_bytes = 0;
    double* VT[2] = {nullptr, nullptr};        // [2*nn] each
    unsigned long long* flags = nullptr;       // [256] arrival epochs written by peers
    int cur = 0;                               // which VT buffer holds the current velocity
    int scur = 0;                              // which sigma plane set is current
    int dcur = 0;                              // which damage plane is current (flips only under BBM)

    nsx::DBuf<double> UM, UT, wind, ocean, tau_wi, tau_a, tau_w, ssh, VTM;
    nsx::DBuf<double> disp;                      // displacement accumulated over the sub-cycle loop (tile / direct paths), [2*nn]
    bool have_tau_wi = false;
    nsx::DBuf<double> sig[2][3], dmg[2];       // ping-pong (tiles recompute neighbours' elements from the old state)
    nsx::DBuf<double> conc, thick, snow, conc_young, h_young, hs_young, thick_myi, conc_myi, ridge_ratio;
    nsx::DBuf<double> depth, drag_ui, drag_ui_young, cohesion, t_heal;
    nsx::DBuf<double> surface, delta_x, shape;   // shape: 6 SoA planes [6*ne]
    nsx::DBuf<double> del_ci_ridge_myi;
    nsx::DBuf<double> emass, ecbu;               // element mass, element C_bu
    nsx::DBuf<double> node_mass, rlmass, cbu, fcor, grad_ssh;
    nsx::DBuf<double> ec_e, contrib;             // direct path: element-space rheology constants, staged contributions
    NsxCreateOptions opt{};
    bool resident = false;                       // state-resident persistent solver (k_resident)
    bool direct = false;                         // L2-resident mesh: element kernel + node kernel instead of the tile kernel
    // resident path
    nsx::DBuf<nsx::ResTile> res_tiles;
    nsx::DBuf<uint16_t> res_n2n;
    nsx::DBuf<uint8_t> res_n2n_deg, halo_move;
    nsx::DBuf<int> halo_slot;
    nsx::DBuf<int2> push_ent_mb;                 // owned node -> (send slot, slot in the holder's mailbox)
    void* mailbox = nullptr;                     // inside the halo window: [2][n_mb] 32-byte entries
    int n_mb = 0;
    nsx::DBuf<unsigned long long> res_time;      // %globaltimer stamps of the last resident launch: start, loop end, end
    unsigned long long* h_time = nullptr;        // pinned copy
    int epoch_bump = 0;                          // exchanges the resident launch of the current graph performs
    nsx::DBuf<double> arena;                     // transfer arena (host numbering): all fields of one upload / download call
    int* h_err = nullptr;                        // pinned copy of the device error word
    nsx::DBuf<int> ow_list;                      // open-water nodes to smooth
    nsx::DBuf<int> ow_count;
    nsx::DBuf<int> check_i; nsx::DBuf<double> check_d;
    // device-side diagnostics / regrid check / forcing time interpolation (allocated on first use)
    nsx::DBuf<double> diag;                      // D_conc, D_thick, D_snow_thick, D_sigma[0], D_sigma[1], D_divergence  [6*ne]
    nsx::DBuf<unsigned long long> regrid_keys;   // min-angle bits, ordered keys of min / max jacobian
    nsx::DBuf<double> forcing[3][2];             // interpolated_data[0..1] of wind, ocean (2 planes) and ssh (1 plane)
    bool forcing_loaded[3][2] = {};
    // thermo().  th holds the device pointer of every field; th_planes owns the thermo-only ones
    nsx::thermo::Arrays th{};
    nsx::DBuf<double> th_planes;
    int n_thermo_launch = 0;
    struct ThermoForcing { nsx::DBuf<double> d[2]; bool loaded[2] = {false, false}; };
    std::map<std::string, std::unique_ptr<ThermoForcing>> th_forcing;    // time slices of the element forcing (ExternalData)

    // ---- halo ----
    std::deque<nsx::PeerLink> peers;             // union of send/recv peers
    nsx::DBuf<int> d_send_src, d_send_dst;       // concatenated push tables over all send peers
    nsx::DBuf<int2> d_send_slot;                 // per send entry: (send slot, slot in the holder's mailbox)
    nsx::DBuf<unsigned long long> d_epoch;       // device-resident exchange counter (graph replay safe)
    nsx::DBuf<unsigned int> d_done;              // block completion counter of k_halo_exchange
    int n_send_total = 0;
    nsx::DBuf<int> push_ptr; nsx::DBuf<int2> push_ent;   // owned node -> (send-peer slot, holder's ghost id)
    nsx::DBuf<int> ow_pair;                      // smoother: [0..32) my open-water bit per send peer, [32..64) active links
    nsx::DBuf<uint8_t> elem_nowrite;             // element written by a boundary tile (mixed direct/tile mode)
    nsx::DBuf<int> halo_err;                     // device error word (timeouts)
    bool halo_ready = false;
    bool halo_local = false;                     // all peers live in this process on this device
    cudaIpcMemHandle_t ipc{};

    // ---- timing ----
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_upd = nullptr;                // start of nsx_update
    NsxTiming timing{};
    bool timing_valid = false;
    bool update_timed = false;
    int n_launch = 0;

    // CUDA graphs of one explicitSolve, indexed by the ping-pong parities at entry (cur + 2*scur + 4*dcur)
    cudaGraphExec_t graph_exec[8] = {};
    int graph_cur_out[8] = {}, graph_scur_out[8] = {}, graph_dcur_out[8] = {};
    int graph_launches[8] = {}, graph_nsub[8] = {};
    bool graph_valid = false;
    bool capturing = false;
};

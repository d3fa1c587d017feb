// nsx_kernels.cuh -- hand-written FP64 CUDA kernels (sm_100a) for the explicit sub-cycled momentum solve.
//
// Layout: every field is a flat SoA plane in HBM.  Element planes are indexed by local element id
// (owned first, ghosts after), node planes by local node id (owned first); nodal 2-vectors are stored
// split [u | v] exactly like the reference (FE.cpp:10331-10332).  Element->node connectivity is three
// int32 planes; the node->element incidence needed for the fixed-order reduction of FE.cpp:10445-10467
// is a column-major ELL table (coalesced per column) of staging-slot ids in ASCENDING element order.
//
// No float atomics anywhere: element kernels write their three nodal stress contributions to a staging
// plane, node kernels subtract them in the reference's order starting from grad_ssh.
#pragma once
#include <climits>
#include "nsx_internal.h"
#include "nsx_mesh.h"

namespace nsx {

constexpr int TPB = 256;

// Device-side bounds checks of the index tables the sub-cycle kernels trust (slot connectivity, incidence codes, halo and
// mailbox slots, push lists).  Compiled in with -DNSX_DEBUG_CHECKS (NSX_DEBUG_CHECKS=1 python -m nextsim_b200.build): a
// violated check writes 3000 + its id into the handle's error word, which nsx_download / nsx_check turn into an error.
// This is the stand-in for compute-sanitizer memcheck, which is closed on the GPU pool this was developed on
// (profiles/r2_debug_checks.txt).
#ifdef NSX_DEBUG_CHECKS
#define NSX_DEV_CHECK(cond, errp, id) do { if (!(cond)) atomicCAS((errp), 0, 3000 + (id)); } while (0)
#else
#define NSX_DEV_CHECK(cond, errp, id) do { } while (0)
#endif

__device__ __forceinline__ double ld_nc(const double* p) { return __ldg(p); }

// ---------------------------------------------------------------------------------------------------
// prep elements  (FE.cpp:10235-10341 minus the nodal scatters, which k_prep_nodes gathers).
// One thread per SLOT of the tile decomposition (own slots + redundant halo slots): geometry and the
// per-step constants of the rheology go to slot space (coalesced for the sub-cycle kernel); the writer
// slot of an element also stores the per-element products the reference keeps as members.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_prep_elements(KParams K, int nslots, const int* __restrict__ slot_elem,
                const int* __restrict__ en0, const int* __restrict__ en1, const int* __restrict__ en2,
                const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ UM,
                const double* __restrict__ conc, const double* __restrict__ thick, const double* __restrict__ snow,
                const double* __restrict__ conc_y, const double* __restrict__ h_y, const double* __restrict__ hs_y,
                const double* __restrict__ depth, const double* __restrict__ ssh,
                const double* __restrict__ cohesion, const double* __restrict__ t_heal,
                double* __restrict__ surface, double* __restrict__ delta_x, double* __restrict__ shape,
                double* __restrict__ emass, double* __restrict__ ecbu,
                double* __restrict__ slot_shape, double* __restrict__ slot_ec, double* __restrict__ ec_e)
{
    int const s = blockIdx.x * blockDim.x + threadIdx.x;
    int const ne = K.ne, nn = K.nn;
    if (s >= nslots) return;
    int e = slot_elem[s];
    if (e == INT_MIN) return;           // pad slot of the even-sized slot space
    bool const own = e >= 0;            // halo slots are stored as ~e
    if (!own) e = ~e;
    int const a = en0[e], b = en1[e], c = en2[e];
    // GmshMesh::vertices(indices, um, 1.)  gmshmesh.cpp:1929-1939
    double const xa = x[a] + 1. * UM[a], ya = y[a] + 1. * UM[a + nn];
    double const xb = x[b] + 1. * UM[b], yb = y[b] + 1. * UM[b + nn];
    double const xc = x[c] + 1. * UM[c], yc = y[c] + 1. * UM[c + nn];

    // sides() FE.cpp:1642-1663 and the integer-truncated mean (quirk Q1, FE.cpp:10239)
    double const s0 = hypot(xb - xa, yb - ya);
    double const s1 = hypot(xc - xb, yc - yb);
    double const s2 = hypot(xc - xa, yc - ya);
    int acc = 0;
    acc = __double2int_rz((double)acc + s0);
    acc = __double2int_rz((double)acc + s1);
    acc = __double2int_rz((double)acc + s2);
    double const dx = (double)(acc / 3);

    // jacobian / measure / shapeCoeff  FE.cpp:1613-1618, 1929-1933, 1951-1964
    double jac = (xb - xa) * (yc - ya);
    jac -= (xc - xa) * (yb - ya);
    double const A = 0.5 * fabs(jac);
    double sc[6];
    sc[0] = (yb - yc) / jac;  sc[1] = (yc - ya) / jac;  sc[2] = (ya - yb) / jac;
    sc[3] = (xc - xb) / jac;  sc[4] = (xa - xc) / jac;  sc[5] = (xb - xa) / jac;
    // slot space keeps FOUR planes (dN0/dx, dN1/dx, dN0/dy, dN1/dy): the gradients of the three shape functions sum to
    // zero, the sub-cycle kernels rebuild the third pair (16 B less per slot and sub-cycle from HBM)
    slot_shape[0 * (size_t)nslots + s] = sc[0];  slot_shape[1 * (size_t)nslots + s] = sc[1];
    slot_shape[2 * (size_t)nslots + s] = sc[3];  slot_shape[3 * (size_t)nslots + s] = sc[4];

    double const cc = conc[e], hh = thick[e];
    double const vol = hh * A;                                 // FE.cpp:10450
    if (K.dynamics_type == NSX_DYN_BBM) {
        bool const ice = !(cc <= 0.1);                         // quirk Q4 (FE.cpp:4146-4151)
        double const expC = exp(K.compaction_param * (1. - cc));
        slot_ec[0 * (size_t)nslots + s] = ice ? expC : 0.;     // 0 marks "no ice" (expC > 0 always)
        slot_ec[1 * (size_t)nslots + s] = pow(hh, K.exp_compression) * K.compression_factor * expC;   // Pmax FE.cpp:4192
        slot_ec[2 * (size_t)nslots + s] = cohesion[e];
        slot_ec[3 * (size_t)nslots + s] = 1. / (dx * K.sqrt_nu_rhoi);   // FE.cpp:4232
        slot_ec[4 * (size_t)nslots + s] = K.dte / t_heal[e] * expC;     // FE.cpp:4256-4257
        slot_ec[5 * (size_t)nslots + s] = vol;
    } else {
        // P = Pstar*exp(-C(1-c)) FE.cpp:10684 ; negative marks thick==0 (quirk Q5, FE.cpp:10656)
        slot_ec[0 * (size_t)nslots + s] = (hh == 0.) ? -1. : K.evp_Pstar * exp(-K.evp_C * (1. - cc));
        slot_ec[1 * (size_t)nslots + s] = vol;
    }
    if (!own) return;

    if (ec_e) {                 // element-space copy of the rheology constants for the direct (small-mesh) path
        int const npl = (K.dynamics_type == NSX_DYN_BBM) ? 6 : 2;
        for (int k = 0; k < npl; ++k) ec_e[(size_t)k * ne + e] = slot_ec[(size_t)k * nslots + s];
    }
    delta_x[e] = dx;
    surface[e] = A;
#pragma unroll
    for (int k = 0; k < 6; ++k) shape[(size_t)k * ne + e] = sc[k];

    // slab mass FE.cpp:10255-10269
    double tc = cc, tt = hh, ts = snow[e];
    if (K.young_ice) { tc += conc_y[e]; tt += h_y[e]; ts += hs_y[e]; }
    emass[e] = (tc > 0.) ? (RHOI * tt + RHOS * ts) / tc : 0.;

    // Lemieux basal stress numerator FE.cpp:10273-10308
    double element_ssh = 0.;
    element_ssh += ssh[a]; element_ssh += ssh[b]; element_ssh += ssh[c];
    element_ssh /= 3.;
    double const depth_eff = fmax(0., element_ssh + fmax(2., depth[e]));
    double critical_h = 0., critical_h_mod = 0.;
    if (K.basal_stress_type == NSX_BASAL_LEMIEUX) {
        double mean_keel_depth = K.k1 * hh;
        mean_keel_depth = fmin(mean_keel_depth, cc * 28.);
        critical_h = cc * depth_eff / K.k1;
        critical_h_mod = mean_keel_depth / K.k1;
    }
    ecbu[e] = K.k2 * fmax(0., critical_h_mod - critical_h) * exp(-K.Cb * (1. - cc));
}

// ---------------------------------------------------------------------------------------------------
// prep nodes  (nodal scatters of FE.cpp:10309-10340 as ordered gathers + FE.cpp:10356-10416)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_prep_nodes(KParams K, const uint8_t* __restrict__ nflags,
             const int* __restrict__ n2e, const int* __restrict__ n2e_deg,
             const int* __restrict__ nec, int nec_w,
             const int* __restrict__ en0, const int* __restrict__ en1, const int* __restrict__ en2,
             const double* __restrict__ surface, const double* __restrict__ emass, const double* __restrict__ ecbu,
             const double* __restrict__ shape, const double* __restrict__ ssh,
             const double* __restrict__ drag_ui, const double* __restrict__ drag_ui_y,
             const double* __restrict__ conc, const double* __restrict__ conc_y,
             const double* __restrict__ wind, const double* __restrict__ lat,
             double* __restrict__ VT, double* __restrict__ VTM,
             double* __restrict__ node_mass, double* __restrict__ rlmass, double* __restrict__ cbu,
             double* __restrict__ fcor, double* __restrict__ grad_ssh, double* __restrict__ tau_a,
             int* __restrict__ ow_list, int* __restrict__ ow_count)
{
    int const n = blockIdx.x * blockDim.x + threadIdx.x;
    int const nn = K.nn, ne = K.ne;
    if (n >= nn) return;
    uint8_t const fl = nflags[n];
    bool const skip_static = (fl & (NF_DIRICHLET | NF_GHOST)) != 0;
    double const g3rd = GRAVITY / 3.;

    double sumA = 0., sumMA = 0., cb = 0., gu = 0., gv = 0.;
    int const deg = n2e_deg[n];
    for (int k = 0; k < deg; ++k) {                       // ascending element id == reference loop order
        int const s = n2e[(size_t)k * nn + n];
        int const i = s / ne;
        int const e = s - i * ne;
        (void)i;
        double const A = surface[e], m = emass[e];
        sumA += A;
        sumMA += m * A;
        cb = fmax(cb, ecbu[e]);
        // FE.cpp:10328 tests the RUNNING nodal mass (this element already added)
        if (!skip_static && sumMA != 0.) {
            double const mgA = m * A * g3rd;
            int const nj0 = en0[e], nj1 = en1[e], nj2 = en2[e];
            double const h0 = ssh[nj0], h1 = ssh[nj1], h2 = ssh[nj2];
            gu -= shape[0 * (size_t)ne + e] * mgA * h0;  gv -= shape[3 * (size_t)ne + e] * mgA * h0;
            gu -= shape[1 * (size_t)ne + e] * mgA * h1;  gv -= shape[4 * (size_t)ne + e] * mgA * h1;
            gu -= shape[2 * (size_t)ne + e] * mgA * h2;  gv -= shape[5 * (size_t)ne + e] * mgA * h2;
        }
    }
    grad_ssh[n] = gu;
    grad_ssh[n + nn] = gv;
    cbu[n] = cb;

    // open-water nodes: zero the velocity (FE.cpp:10366-10370) and remember them for the smoother
    if (sumMA == 0.) {
        VT[n] = 0.;
        VT[n + nn] = 0.;
        if (n < K.ndof && !(fl & NF_DIRICHLET)) {
            int const slot = atomicAdd(ow_count, 1);
            ow_list[slot] = n;
        }
    }

    // atmospheric drag, surface-weighted over bamg's NodalElementConnectivity order (FE.cpp:10374-10394)
    double drag = 0., surf = 0.;
    for (int j = 0; j < nec_w; ++j) {
        int const e = nec[(size_t)j * nn + n];
        if (e < 0) continue;
        double dragp = drag_ui[e];
        if (K.young_ice) {
            double const c0 = conc[e], c1 = conc_y[e];
            if (c0 + c1 > 0.) dragp = (drag_ui[e] * c0 + drag_ui_y[e] * c1) / (c0 + c1);
        }
        drag += dragp * surface[e];
        surf += surface[e];
    }
    double const wu = wind[n], wv = wind[n + nn];
    drag *= RHOA * hypot(wu, wv) / surf;
    tau_a[n] = drag * wu;
    tau_a[n + nn] = drag * wv;

    fcor[n] = 2 * OMEGA * sin(lat[n] * PI_ / 180.);

    double rl = 1. / sumA;                                  // FE.cpp:10400-10402
    node_mass[n] = sumMA * rl;
    rl *= 3.;
    rlmass[n] = rl;

    VTM[n] = VT[n];
    VTM[n + nn] = VT[n + nn];
}

// ---------------------------------------------------------------------------------------------------
// The sub-cycle kernel: ONE launch per sub-cycle, one CTA per tile, tile data staged by TMA.
//   stage    one thread issues cp.async.bulk (TMA, SASS UBLKCP) copies of every contiguous piece of the
//            tile -- slot planes (connectivity, shape coefficients, rheology constants), the writer slots'
//            sigma/damage, the owned nodes' planes, incidence table, halo lists -- into shared memory; all
//            threads wait on one mbarrier.  Sources only need 8-byte alignment: each copy starts at the
//            enclosing 16-byte boundary and the consumer applies the resulting element shift.
//   phase 0  gather the velocities of the tile's halo nodes (the only irregular read besides halo sigma)
//   phase 1  per slot: strain rate from the staged velocities, BBM (FE.cpp:4137-4260) or EVP/mEVP
//            (FE.cpp:10649-10726) stress update, write sigma/damage (writer slots only; ping-pong planes, so
//            tiles recomputing a neighbour's element read the old state), and leave the three nodal
//            contributions V*(sigma.grad N_i) (FE.cpp:10464-10465) in shared memory (over the shape planes)
//   phase 2  per owned node: subtract the contributions in ASCENDING reference element order starting from
//            grad_ssh (no float atomics, FE.cpp:10445-10467), implicit drag/Coriolis/basal 2x2 solve
//            (FE.cpp:10472-10529), write VT into the other ping-pong buffer, move the mesh (10539-10553);
//            plus the lagged mesh move of this tile's share of the ghost nodes.
// HBM traffic per element-sub-cycle ~ 104 B slot constants + 32 B sigma/d read (x ~1.1 redundancy) + 32 B
// written + ~95 B of nodal planes = ~280 B (SURVEY 8(d) algorithmic floor: 264 B).
// ---------------------------------------------------------------------------------------------------
#ifndef NSX_SUB_TPB
#define NSX_SUB_TPB 768
#endif
#ifndef NSX_SUB_MINB
#define NSX_SUB_MINB 2
#endif
#ifndef NSX_SUB_STAGES
#define NSX_SUB_STAGES 2
#endif
constexpr int SUB_TPB = NSX_SUB_TPB;
constexpr int SUB_STAGES = NSX_SUB_STAGES;     // shared-memory stages of the tile pipeline

// ---- fast FP64 reciprocal / square root for the sub-cycle kernel ----
// Hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 bits) + two Newton steps + one residual correction: <= 2 ulp,
// a quarter of the instructions and of the dependent latency of the IEEE sequences.  Parity tolerance is 1e-9
// relative L2, so the last-bit difference to the reference's libm/IEEE results is immaterial; arguments outside
// a comfortable exponent range (zeros, denormals, infinities) take the exact slow path.
__device__ __forceinline__ bool mid_range(double x)
{
    double const ax = fabs(x);
    return ax > 1e-280 && ax < 1e280;
}
__device__ __forceinline__ double fast_div(double a, double b)
{
    if (!mid_range(b)) return a / b;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    double const q = a * r;
    return fma(r, fma(-b, q, a), q);
}
__device__ __forceinline__ double fast_sqrt(double x)
{
    if (!mid_range(x) || x < 0.) return sqrt(x);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double const h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    double const sq = x * y;
    return fma(fma(-sq, sq, x), 0.5 * y, sq);
}
__device__ __forceinline__ double fast_hypot(double a, double b)
{
    double const q = fma(a, a, b * b);
    if (!mid_range(q)) return hypot(a, b);
    return fast_sqrt(q);
}

__device__ __forceinline__ double pow_relax(double q, KParams const& K)
{
    switch (K.relax_int_pow) {
        case 0: return 1.;
        case 1: return q;
        case 2: return q * q;
        case 3: return q * q * q;
        case 4: { double const q2 = q * q; return q2 * q2; }
        default: return pow(q, K.exp_relax_m1);
    }
}

// ---- TMA / mbarrier primitives (PTX ISA 8.x, sm_90+) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// element shift of an 8-byte-or-less aligned source inside its enclosing 16-byte granule
__device__ __forceinline__ int stage_shift(const void* src, int esz) { return (int)(((uintptr_t)src & 15) / esz); }
// MODE 0: returns the bytes the copy will transfer; MODE 1: issues it
template <int MODE>
__device__ __forceinline__ uint32_t stage(void* dst16, const void* src, int n, int esz, uint64_t* bar, int& ctr)
{
    if (n <= 0) return 0;
    uint32_t const lead = (uint32_t)((uintptr_t)src & 15);
    uint32_t const bytes = (lead + (uint32_t)n * esz + 15u) & ~15u;
    if (MODE) { bulk_g2s(dst16, (const void*)((uintptr_t)src - lead), bytes, bar); ++ctr; }
    return bytes;
}

// byte offsets of the staging buffers inside dynamic shared memory (all multiples of 16) and plane strides
struct SmemLayout {
    int bar, conn, shape, ec, sig, dmg, node, su, sv, hsig, inc, fl, total;
    int msp, mop, mtp, mhs;     // slot / own-slot / node / halo-slot plane strides (elements)
};
enum { NP_GSU, NP_GSV, NP_MASS, NP_RL, NP_CBU, NP_FCOR, NP_TAU, NP_TAV, NP_OCU, NP_OCV, NP_DSU, NP_DSV,
       NP_VMU, NP_VMV, NP_TWU, NP_TWV, NP_COUNT };

struct HaloArgs {
    int n_total;                        // send entries over all peers
    int n_peers;                        // peers I send to
    int peer_begin[33];                 // entry ranges per send peer
    double* peer_vt[32];                // holder's VT buffer (current parity)
    int peer_nn[32];
    int peer_link[32];                  // send peer -> its index in the link lists below
    // every neighbour (send or receive relation) is signalled AND waited for, so the two ranks of a pair never drift
    // by more than one exchange: a push can only land in a buffer its holder has finished reading
    int n_link;
    unsigned long long* link_flag[32];  // the neighbour's flag slot for me
    int link_rank[32];                  // my flag slot that neighbour writes
    int sync;                           // 0: stream-ordered group on one device, no flags
};
// Flag word = exchange epoch; bit 63 is the "my send list to you contains open-water nodes" marker of the smoother
// (k_ow_sweep_exchange).  Every waiter compares the masked value.
constexpr unsigned long long FLAG_OW = 0x8000000000000000ULL;

// Publishes `word` in the flag slot of every neighbour selected by `mask` and waits until each of them has
// published an epoch >= `epoch` (bounded spin).  Called by ONE block after its pushes; thread i serves link i.
// Returns (to thread i) the last flag word read from neighbour i, 0 if not selected.
__device__ __forceinline__ unsigned long long flag_signal_wait(HaloArgs const& a, int i, bool selected, unsigned long long word,
                                                               unsigned long long epoch, const unsigned long long* my_flags,
                                                               long long max_spins, int* err)
{
    unsigned long long v = 0ULL;
    if (i < a.n_link && selected) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.link_flag[i]), "l"(word) : "memory");
        const unsigned long long* f = my_flags + a.link_rank[i];
        long long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if ((v & ~FLAG_OW) >= epoch) break;
            if (*((volatile int*)err)) break;        // an earlier exchange already timed out: do not stall again
            if (++spins > max_spins) { atomicExch(err, 1 + a.link_rank[i]); break; }
            __nanosleep(64);
        }
    }
    return v;
}

// ---- mailbox exchange (synchronisation carried by the data; see the k_resident header) ----
constexpr int MB_MAX_PEERS = 16;
// mailbox entry of one node: two 16-byte {value, tag} pairs, each written / read as ONE vector transaction
struct __align__(16) MbEntry { double u; unsigned long long tu; double v; unsigned long long tv; };
static_assert(sizeof(MbEntry) == 32, "mailbox entry");

__device__ __forceinline__ void mb_store(MbEntry* e, double u, double v, unsigned long long tag)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(&e->u), "l"(__double_as_longlong(u)), "l"(tag) : "memory");
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(&e->v), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
// polls until both halves carry `tag`; bounded (sets *err and gives up, also as soon as another waiter has timed out)
__device__ __forceinline__ void mb_wait(const MbEntry* e, unsigned long long tag, double& u, double& v, int* err, int who)
{
    unsigned long long a = 0, ta = 0, b = 0, tb = 0;
    long long spins = 0;
    for (;;) {
        if (ta != tag) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(ta) : "l"(&e->u) : "memory");
        if (tb != tag) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(b), "=l"(tb) : "l"(&e->v) : "memory");
        if (ta == tag && tb == tag) break;
        if ((++spins & 1023) == 0 && (*((volatile int*)err) || spins > (1LL << 22))) { atomicCAS(err, 0, who); break; }
    }
    u = __longlong_as_double((long long)a);
    v = __longlong_as_double((long long)b);
}

// mailbox exchange of the tile / direct paths across GPUs (one process per GPU): sent nodes go into the holders' mailboxes
// from the kernel that computes them, the next sub-cycle's kernels read ghost velocities from this rank's mailbox (polling
// the tag) -- no flag, no fence, no second stream.  Mailbox slot of ghost g = g - ndof; THREE buffers, exchange ex uses
// buffer ex % 3: a neighbour can be at most one launch ahead (its launch for exchange e+1 needs this rank's pushes of e,
// issued by a launch that started after the previous one had finished), so while some thread here still reads exchange
// e-1 the neighbour writes at most e+1 -- never the buffer of e-1.  (Two buffers would need the tile-level symmetry the
// resident kernel has; the ghost-move shares and the direct kernels do not have it.)
struct MbExchange {
    int on;                                     // 0: no neighbour ranks / in-process group (stream-ordered exchange)
    int ex;                                     // exchange index inside the model step (1-based); tag = *epoch_ctr + ex
    MbEntry* mb; int n_mb;                      // this rank's mailbox [2][n_mb]
    MbEntry* peer_mb[MB_MAX_PEERS]; int peer_nmb[MB_MAX_PEERS];
    const int* push_ptr; const int2* push_ent;  // owned node -> (send slot, slot in the holder's mailbox)
    const unsigned long long* epoch_ctr; int* err;
};
__device__ __forceinline__ void mbx_push(MbExchange const& X, int n, double un, double vn)
{
    int const q1 = X.push_ptr[n + 1];
    int q = X.push_ptr[n];
    if (q == q1) return;
    unsigned long long const tag = *X.epoch_ctr + (unsigned long long)X.ex;
    for (; q < q1; ++q) {
        int2 const pe = X.push_ent[q];
        NSX_DEV_CHECK(pe.x >= 0 && pe.x < MB_MAX_PEERS && pe.y >= 0 && pe.y < X.peer_nmb[pe.x], X.err, 21);
        mb_store(X.peer_mb[pe.x] + (size_t)(X.ex % 3) * X.peer_nmb[pe.x] + pe.y, un, vn, tag);
    }
}
__device__ __forceinline__ void mbx_import(MbExchange const& X, int slot, double& u, double& v)
{
    NSX_DEV_CHECK(slot >= 0 && slot < X.n_mb, X.err, 22);
    mb_wait(X.mb + (size_t)(X.ex % 3) * X.n_mb + slot, *X.epoch_ctr + (unsigned long long)X.ex, u, v, X.err, 500 + slot % 400);
}
// velocity of node g as the sub-cycle kernels of exchange X.ex read it: a ghost node comes from the mailbox of the PREVIOUS
// exchange (its owner pushed it from the previous sub-cycle's launch on its GPU; polling covers a peer that lags), any
// other node -- and every node in the first sub-cycle -- from the velocity buffer
__device__ __forceinline__ void mbx_velocity(MbExchange const& X, int g, int ndof, int nn, const double* VT, double& u, double& v)
{
    if (X.on && X.ex > 1 && g >= ndof) {
        int const slot = g - ndof, pe = X.ex - 1;
        NSX_DEV_CHECK(slot < X.n_mb, X.err, 23);
        mb_wait(X.mb + (size_t)(pe % 3) * X.n_mb + slot, *X.epoch_ctr + (unsigned long long)pe, u, v, X.err, 500 + slot % 400);
    } else { u = VT[g]; v = VT[g + nn]; }
}

struct SubArgs {
    const TileDesc* tiles; const int* tile_order; int tile_base;
    const int* halo_nodes; const int* halo_elems; const unsigned long long* slot_conn;
    const double* slot_shape; const double* slot_ec; int nslots; const uint16_t* inc;
    const double* s0i; const double* s1i; const double* s2i; const double* di;
    double* s0o; double* s1o; double* s2o; double* dmo;
    const uint8_t* nflags; const double* grad_ssh; const double* node_mass; const double* rlmass;
    const double* cbu; const double* fcor; const double* tau_a; const double* tau_wi; const double* ocean;
    const double* VTM; const double* VTc; double* VTn; double* disp;
    int move_mesh, lag_ghost_move, n_tiles;
    int np[NP_COUNT];           // node plane -> staging slot (compacted: only the planes this configuration reads)
    // fused ghost exchange of the boundary launch (multi-GPU): phase 2 stores sent nodes straight into the holders'
    // ghost slots over NVLink; the last CTA publishes the epoch and waits for the owners of this rank's ghosts
    int fuse_halo;
    const int* push_ptr; const int2* push_ent;
    const unsigned long long* my_flags; unsigned long long* epoch_ctr; unsigned int* done_ctr; int* halo_err;
    HaloArgs H;
    MbExchange X;
    SmemLayout L;
};

// Issues the TMA copies of copy-group `g` (0..3, one group per producer warp so that the four warps issue in
// parallel) and returns the bytes they will deliver.  The caller posts arrive.expect_tx AFTERWARDS: the
// transaction count of an mbarrier may go negative, and the phase cannot complete before that arrival.
template <int BBM>
__device__ __forceinline__ uint32_t stage_tile(int g, KParams const& K, SubArgs const& A, TileDesc const& td, unsigned char* sm, uint64_t* bar)
{
    SmemLayout const& L = A.L;
    int const nn = K.nn;
    int const nsl = td.n_own_slots + td.n_halo_slots;
    size_t const NS = (size_t)A.nslots;
    size_t const s0 = (size_t)td.slot_begin;
    int const nb = td.node_begin, no = td.n_own;
    uint32_t tx = 0;
    int ctr = 0;
    auto node_plane = [&](int p, const double* src) { return stage<1>(sm + L.node + (size_t)A.np[p] * L.mtp * 8, src + nb, no, 8, bar, ctr); };
    if (g == 0) {
        tx += stage<1>(sm + L.conn, A.slot_conn + s0, nsl, 8, bar, ctr);
        // four staged shape planes land in planes 0, 1 (d/dx of N0, N1) and 3, 4 (d/dy); planes 2 and 5 only ever hold
        // contributions
#pragma unroll
        for (int p = 0; p < 4; ++p)
            tx += stage<1>(sm + L.shape + (size_t)(p < 2 ? p : p + 1) * L.msp * 8, A.slot_shape + p * NS + s0, nsl, 8, bar, ctr);
        tx += stage<1>(sm + L.su, A.VTc + nb, no, 8, bar, ctr);
        tx += stage<1>(sm + L.sv, A.VTc + nn + nb, no, 8, bar, ctr);
    } else if (g == 1) {
#pragma unroll
        for (int p = 0; p < (BBM ? 6 : 2); ++p)
            tx += stage<1>(sm + L.ec + (size_t)p * L.msp * 8, A.slot_ec + p * NS + s0, nsl, 8, bar, ctr);
        tx += stage<1>(sm + L.sig + (size_t)0 * L.mop * 8, A.s0i + td.elem_begin, td.n_own_slots, 8, bar, ctr);
        tx += stage<1>(sm + L.sig + (size_t)1 * L.mop * 8, A.s1i + td.elem_begin, td.n_own_slots, 8, bar, ctr);
        tx += stage<1>(sm + L.sig + (size_t)2 * L.mop * 8, A.s2i + td.elem_begin, td.n_own_slots, 8, bar, ctr);
        if (BBM) tx += stage<1>(sm + L.dmg, A.di + td.elem_begin, td.n_own_slots, 8, bar, ctr);
    } else if (g == 2) {
        tx += node_plane(NP_GSU, A.grad_ssh);   tx += node_plane(NP_GSV, A.grad_ssh + nn);
        tx += node_plane(NP_MASS, A.node_mass); tx += node_plane(NP_RL, A.rlmass);
        tx += node_plane(NP_CBU, A.cbu);        tx += node_plane(NP_FCOR, A.fcor);
        tx += node_plane(NP_TAU, A.tau_a);      tx += node_plane(NP_TAV, A.tau_a + nn);
        tx += node_plane(NP_OCU, A.ocean);      tx += node_plane(NP_OCV, A.ocean + nn);
    } else {
        if (A.move_mesh) { tx += node_plane(NP_DSU, A.disp); tx += node_plane(NP_DSV, A.disp + nn); }
        if (K.dynamics_type == NSX_DYN_MEVP) { tx += node_plane(NP_VMU, A.VTM); tx += node_plane(NP_VMV, A.VTM + nn); }
        if (A.tau_wi) { tx += node_plane(NP_TWU, A.tau_wi); tx += node_plane(NP_TWV, A.tau_wi + nn); }
        tx += stage<1>(sm + L.inc, A.inc + td.inc_off, td.inc_w * td.n_own, 2, bar, ctr);
        tx += stage<1>(sm + L.fl, A.nflags + nb, no, 1, bar, ctr);
    }
    return tx;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int SUB_PROD = 128;                   // producer threads (4 warps): TMA issue + irregular gathers
constexpr int SUB_CONS = SUB_TPB - SUB_PROD;    // consumer threads
#ifndef NSX_SUB_GROUPS
#define NSX_SUB_GROUPS 1
#endif
constexpr int SUB_GROUPS = NSX_SUB_GROUPS;      // consumer groups, each working on its own tile (latency chains overlap)
constexpr int SUB_GS = SUB_CONS / SUB_GROUPS;   // threads per consumer group
static_assert(SUB_GS % 32 == 0 && SUB_GS * SUB_GROUPS == SUB_CONS, "consumer groups must be whole warps");
static_assert(SUB_STAGES > SUB_GROUPS || SUB_GROUPS == 1, "need one more stage than tiles in compute");
__device__ __forceinline__ void cons_sync(int group)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(SUB_GS) : "memory");
}

__device__ __forceinline__ void p2_sync()       // barrier over the consumer threads of group 0 (fused halo epilogue)
{
    asm volatile("bar.sync 1, %0;" ::"n"(SUB_GS) : "memory");
}

// Persistent, warp-specialised, double-buffered: CTA b works on tiles b, b+grid, b+2*grid, ... of its launch
// range.  The producer warp streams tile t+1 into the other stage while the consumer warps compute tile t
// (full[s]: TMA transaction barrier, empty[s]: one arrival per consumer warp when the stage may be refilled).
#ifndef NSX_SUB_CTAS_PER_SM
#define NSX_SUB_CTAS_PER_SM 1
#endif
constexpr int SUB_CTAS_PER_SM = NSX_SUB_CTAS_PER_SM;      // persistent CTAs per SM (each with its own stage ring)

template <int BBM>
__global__ void __launch_bounds__(SUB_TPB, NSX_SUB_CTAS_PER_SM)
k_subcycle(KParams K, SubArgs A)
{
    extern __shared__ __align__(128) unsigned char sm_all[];
    SmemLayout const& L = A.L;
    int const nn = K.nn;
    int const tid = threadIdx.x;
    uint64_t* const full = (uint64_t*)sm_all;
    uint64_t* const empty = full + SUB_STAGES;
    unsigned char* const stage0 = sm_all + 128;     // header: barriers [0,96), fused-halo flag at 120
    if (tid == 0) {
        for (int q = 0; q < SUB_STAGES; ++q) {
            mbar_init(full + q, SUB_PROD / 32);       // one arrival (with its TMA bytes) per producer warp
            mbar_init(empty + q, SUB_GS / 32);        // one arrival per consumer warp of the group that reads the stage
        }
    }
    __syncthreads();
    int const n_my = (A.n_tiles > (int)blockIdx.x) ? (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid >= SUB_CONS) {
        // ---- producer warps: thread 0 streams the contiguous pieces with TMA; all producer threads gather the
        // irregular ones (halo node velocities, halo slots' sigma/damage), one entry each, one tile ahead ----
        int const p = tid - SUB_CONS;
        for (int it = 0; it < n_my; ++it) {
            int const s = it % SUB_STAGES;
            int const tix = A.tile_base + (int)blockIdx.x + it * (int)gridDim.x;
            TileDesc const td = A.tiles[A.tile_order ? A.tile_order[tix] : tix];
            unsigned char* const sm = stage0 + (size_t)s * L.total;
            if (it >= SUB_STAGES) mbar_wait(empty + s, ((it / SUB_STAGES) - 1) & 1);
            uint32_t tx = 0;
            if (p == 0) *(TileDesc*)(sm + L.bar) = td;              // consumers read the descriptor from the stage
            if ((p & 31) == 0) tx = stage_tile<BBM>(p >> 5, K, A, td, sm, full + s);
            double* const su = (double*)(sm + L.su) + stage_shift(A.VTc + td.node_begin, 8);
            double* const sv = (double*)(sm + L.sv) + stage_shift(A.VTc + nn + td.node_begin, 8);
            double* const hsg = (double*)(sm + L.hsig);
            int const nmax = max(td.n_halo, td.n_halo_slots);
            for (int j = p; j < nmax; j += SUB_PROD) {
                bool const hn_ok = j < td.n_halo, he_ok = j < td.n_halo_slots;
                int const g = hn_ok ? A.halo_nodes[td.halo_off + j] : 0;
                int const e = he_ok ? A.halo_elems[td.halo_elem_off + j] : 0;
                double u, v;
                mbx_velocity(A.X, g, K.ndof, nn, A.VTc, u, v);
                double const a0 = A.s0i[e], a1 = A.s1i[e], a2 = A.s2i[e];
                double const ad = BBM ? A.di[e] : 0.;
                if (hn_ok) { su[td.n_own + HALO_GAP + j] = u; sv[td.n_own + HALO_GAP + j] = v; }
                if (he_ok) { hsg[j] = a0; hsg[L.mhs + j] = a1; hsg[2 * L.mhs + j] = a2; if (BBM) hsg[3 * L.mhs + j] = ad; }
            }
            __syncwarp();
            if ((p & 31) == 0) mbar_expect_tx(full + s, tx);       // arrival of this warp: its gathers are done
        }
        return;
    }

    // ---- consumers ----
    int const grp = tid / SUB_GS;               // consumer group; it takes tiles grp, grp + SUB_GROUPS, ...
    int const gtid = tid - grp * SUB_GS;
    int const gstride = SUB_GS;
    for (int it = grp; it < n_my; it += SUB_GROUPS) {
    int const s = it % SUB_STAGES;
    unsigned char* const sm = stage0 + (size_t)s * L.total;
    mbar_wait(full + s, (it / SUB_STAGES) & 1);
    TileDesc const td = *(const TileDesc*)(sm + L.bar);

    // shifted views of the staged planes
    int const nb = td.node_begin;
    int const sh_slot = stage_shift(A.slot_conn + td.slot_begin, 8);      // nslots is even: same phase for every slot plane
    int const sh_el = stage_shift(A.s0i + td.elem_begin, 8);
    int const sh_nu = stage_shift(A.VTc + nb, 8), sh_nv = stage_shift(A.VTc + nn + nb, 8);
    const unsigned long long* const conn = (const unsigned long long*)(sm + L.conn) + sh_slot;
    double* const shp = (double*)(sm + L.shape) + sh_slot;                // also the contribution planes
    const double* const ecp = (const double*)(sm + L.ec) + sh_slot;
    const double* const sgp = (const double*)(sm + L.sig) + sh_el;
    const double* const dgp = (const double*)(sm + L.dmg) + sh_el;
    double* const su = (double*)(sm + L.su) + sh_nu;
    double* const sv = (double*)(sm + L.sv) + sh_nv;
    const double* const hsg = (const double*)(sm + L.hsig);       // halo slots' sigma/damage, gathered by the producer warp
    int const MHS = L.mhs;
    const uint16_t* const incp = (const uint16_t*)(sm + L.inc) + stage_shift(A.inc + td.inc_off, 2);
    const uint8_t* const flp = (const uint8_t*)(sm + L.fl) + stage_shift(A.nflags + nb, 1);
    int const MSP = L.msp, MOP = L.mop, MTP = L.mtp;

    // ---- phase 1 ----
    int const nsl = td.n_own_slots + td.n_halo_slots;
    for (int k = gtid; k < nsl; k += gstride) {
        bool const own = k < td.n_own_slots;
        int const e = td.elem_begin + k;            // meaningful for writer slots only
        int const hk = k - td.n_own_slots;
        unsigned long long const pc = conn[k];
        int const la = (int)(pc & 0xFFFF), lb = (int)((pc >> 16) & 0xFFFF), lc = (int)((pc >> 32) & 0xFFFF);
        NSX_DEV_CHECK(la < td.n_own + HALO_GAP + td.n_halo && lb < td.n_own + HALO_GAP + td.n_halo && lc < td.n_own + HALO_GAP + td.n_halo &&
                      k < MSP && (!own || e < K.ne), A.halo_err, 11);
        double const dx0 = shp[k], dx1 = shp[MSP + k], dy0 = shp[3 * MSP + k], dy1 = shp[4 * MSP + k];
        double const dx2 = -(dx0 + dx1), dy2 = -(dy0 + dy1);      // sum of the shape-function gradients is zero
        double const c0 = ecp[k];
        double s0, s1, s2, vol;
        if (BBM) {
            double const expC = c0;
            vol = ecp[5 * MSP + k];
            double d;
            if (expC == 0.) {                   // conc <= 0.1 : no ice (FE.cpp:4151-4159)
                s0 = s1 = s2 = 0.;
                d = 0.;
            } else {
                double const ua = su[la], va = sv[la], ub = su[lb], vb = sv[lb], uc = su[lc], vc = sv[lc];
                // epsilon_veloc = B0T * u  (FE.cpp:4167-4176; B0T rows: [dxN 0], [0 dyN], [dyN dxN])
                double e0 = dx0 * ua; e0 += dx1 * ub; e0 += dx2 * uc;
                double e1 = dy0 * va; e1 += dy1 * vb; e1 += dy2 * vc;
                double e2 = dy0 * ua; e2 += dx0 * va; e2 += dy1 * ub; e2 += dx1 * vb; e2 += dy2 * uc; e2 += dx2 * vc;
                if (own) { s0 = sgp[k]; s1 = sgp[MOP + k]; s2 = sgp[2 * MOP + k]; d = dgp[k]; }
                else     { s0 = hsg[hk]; s1 = hsg[MHS + hk]; s2 = hsg[2 * MHS + hk]; d = hsg[3 * MHS + hk]; }
                double const dt = K.dte;
                double sigma_n = (s0 + s1) * 0.5;
                double const omd = 1. - d;
                double const time_viscous = K.lambda0 * pow_relax(omd * expC, K);
                double tildeP = 0.;
                if (sigma_n < 0.) {
                    double const Pmax = ecp[MSP + k];
                    tildeP = fmin(1., fast_div(-Pmax, sigma_n));
                }
                double const mult = fmin(1. - 1e-12, fast_div(time_viscous, time_viscous + dt * (1. - tildeP)));
                double const elasticity = K.young * omd * expC;
                double const dtE = dt * elasticity;
                s0 += dtE * K.D00 * e0;  s0 += dtE * K.D01 * e1;  s0 *= mult;
                s1 += dtE * K.D01 * e0;  s1 += dtE * K.D00 * e1;  s1 *= mult;
                s2 += dtE * K.D22 * e2;                           s2 *= mult;
                double const sigma_s = fast_hypot((s0 - s1) * 0.5, s2);
                sigma_n = (s0 + s1) * 0.5;
                double dcrit;
                if (sigma_n < -K.compr_strength) dcrit = fast_div(-K.compr_strength, sigma_n);
                else dcrit = fast_div(ecp[2 * MSP + k], sigma_s + K.tan_phi * sigma_n);
                if ((0. < dcrit) && (dcrit < 1.)) {
                    double const rtd = fast_sqrt(elasticity) * ecp[3 * MSP + k];
                    double const f = (1. - dcrit) * dt * rtd;
                    d += omd * f;
                    s0 -= s0 * f;  s1 -= s1 * f;  s2 -= s2 * f;
                }
                d = fmax(0., d - ecp[4 * MSP + k]);
            }
            if (own) A.dmo[e] = d;
        } else {
            double const Pp = c0;
            vol = ecp[MSP + k];
            if (Pp < 0.) {                      // thick == 0 (FE.cpp:10656-10662)
                s0 = s1 = s2 = 0.;
            } else {
                double const ua = su[la], va = sv[la], ub = su[lb], vb = sv[lb], uc = su[lc], vc = sv[lc];
                double eps11 = dx0 * ua; eps11 += dx1 * ub; eps11 += dx2 * uc;
                double eps22 = dy0 * va; eps22 += dy1 * vb; eps22 += dy2 * vc;
                double eps12 = 0.5 * (dx0 * va + dy0 * ua); eps12 += 0.5 * (dx1 * vb + dy1 * ub); eps12 += 0.5 * (dx2 * vc + dy2 * uc);
                double const eps1 = eps11 + eps22, eps2 = eps11 - eps22;
                double const delta = fast_sqrt(eps1 * eps1 + (eps2 * eps2 + 4 * eps12 * eps12) * K.re2);
                double const zeta = fast_div(Pp, delta + K.evp_dmin);
                if (own) { s0 = sgp[k]; s1 = sgp[MOP + k]; s2 = sgp[2 * MOP + k]; }
                else     { s0 = hsg[hk]; s1 = hsg[MHS + hk]; s2 = hsg[2 * MHS + hk]; }
                double sigma1 = s0 + s1, sigma2 = s0 - s1;
                sigma1 += K.ralpha1 * (zeta * (eps1 - delta) - sigma1);
                sigma2 += K.ralpha2 * (zeta * eps2 * K.re2 - sigma2);
                s2 += K.ralpha2 * (zeta * eps12 * K.re2 - s2);
                s0 = 0.5 * (sigma1 + sigma2);
                s1 = 0.5 * (sigma1 - sigma2);
            }
        }
        if (own) { A.s0o[e] = s0; A.s1o[e] = s1; A.s2o[e] = s2; }
        // nodal contributions V*(sigma . grad N_i) (FE.cpp:10464-10465) overwrite this slot's shape coefficients
        shp[0 * MSP + k] = vol * (s0 * dx0 + s2 * dy0);
        shp[1 * MSP + k] = vol * (s0 * dx1 + s2 * dy1);
        shp[2 * MSP + k] = vol * (s0 * dx2 + s2 * dy2);
        shp[3 * MSP + k] = vol * (s2 * dx0 + s1 * dy0);
        shp[4 * MSP + k] = vol * (s2 * dx1 + s1 * dy1);
        shp[5 * MSP + k] = vol * (s2 * dx2 + s1 * dy2);
    }
    cons_sync(grp);

    // ---- phase 2 ----
    const double* const npl = (const double*)(sm + L.node);
    int const shs = stage_shift(A.node_mass + nb, 8);       // scalar node planes and u halves: same phase as nb
    for (int j = gtid; j < td.n_own; j += gstride) {
        int const n = nb + j;
        uint8_t const fl = flp[j];
        double const uice = su[j], vice = sv[j];
        double un = uice, vn = vice;
        double const nm = npl[A.np[NP_MASS] * MTP + shs + j];
        if (!(fl & NF_DIRICHLET) && nm != 0.) {
            double gu = npl[A.np[NP_GSU] * MTP + shs + j], gv = npl[A.np[NP_GSV] * MTP + sh_nv + j];
            const uint16_t* ip = incp + j;
            for (int c = 0; c < td.inc_w; ++c) {
                unsigned const code = ip[c * td.n_own];
                if (code == 0xFFFFu) break;
                NSX_DEV_CHECK((int)code < 3 * MSP && (int)(code % (unsigned)MSP) < nsl, A.halo_err, 12);
                gu -= shp[code];
                gv -= shp[code + 3 * MSP];
            }
            double dtep = K.dte, delu = 0., delv = 0.;
            if (K.dynamics_type == NSX_DYN_MEVP) {
                delu = (npl[A.np[NP_VMU] * MTP + shs + j] - uice) * K.mevp_rb;
                delv = (npl[A.np[NP_VMV] * MTP + sh_nv + j] - vice) * K.mevp_rb;
                dtep = K.dte_mevp;
            }
            double const dte_over_mass = fast_div(dtep, fmax(K.min_m, nm));
            double const ou = npl[A.np[NP_OCU] * MTP + shs + j], ov = npl[A.np[NP_OCV] * MTP + sh_nv + j];
            double const c_prime = K.rhow_cdw * fast_hypot(ou - uice, ov - vice);
            double const tau_b = npl[A.np[NP_CBU] * MTP + shs + j] * fast_div(1., fast_hypot(uice, vice) + K.u0);
            double const sin_s = (fl & NF_LATNEG) ? -K.sin_ota_abs : K.sin_ota_abs;   // std::copysign(sin, lat[i])
            double const alpha = 1. + dte_over_mass * (c_prime * K.cos_ota + tau_b);
            double const beta = dtep * npl[A.np[NP_FCOR] * MTP + shs + j] + dte_over_mass * c_prime * sin_s;
            double const rdenom = fast_div(1., alpha * alpha + beta * beta);
            double tau_x = npl[A.np[NP_TAU] * MTP + shs + j], tau_y = npl[A.np[NP_TAV] * MTP + sh_nv + j];
            if (A.tau_wi) { tau_x = tau_x + npl[A.np[NP_TWU] * MTP + shs + j]; tau_y = tau_y + npl[A.np[NP_TWV] * MTP + sh_nv + j]; }
            tau_x = tau_x + c_prime * (ou * K.cos_ota - ov * sin_s);
            tau_y = tau_y + c_prime * (ov * K.cos_ota + ou * sin_s);
            double const rl = npl[A.np[NP_RL] * MTP + shs + j];
            double const grad_x = gu * rl, grad_y = gv * rl;
            un = alpha * uice + beta * vice + dte_over_mass * (alpha * (grad_x + tau_x) + beta * (grad_y + tau_y)) + alpha * delu + beta * delv;
            un *= rdenom;
            vn = alpha * vice - beta * uice + dte_over_mass * (alpha * (grad_y + tau_y) - beta * (grad_x + tau_x)) + alpha * delv - beta * delu;
            vn *= rdenom;
        }
        A.VTn[n] = un;
        A.VTn[n + nn] = vn;
        // updateGhosts (FE.cpp:13963-13996): straight into the holders' mailboxes.  Only boundary tiles own sent nodes; the
        // others must not even look at the push lists (a global load at the tail of phase 2 delays the stage hand-over)
        if (A.X.on && td.boundary) mbx_push(A.X, n, un, vn);
        if (A.fuse_halo) {                      // flag protocol of the in-kernel boundary launch (kept for reference runs)
            int const q1 = A.push_ptr[n + 1];
            for (int q = A.push_ptr[n]; q < q1; ++q) {
                int2 const pe = A.push_ent[q];
                NSX_DEV_CHECK(pe.x >= 0 && pe.x < A.H.n_peers && pe.y >= 0 && pe.y < A.H.peer_nn[pe.x], A.halo_err, 13);
                double* const dst = A.H.peer_vt[pe.x];
                dst[pe.y] = un;
                dst[pe.y + A.H.peer_nn[pe.x]] = vn;
            }
        }
        if (A.move_mesh) {      // M_UM and M_UT both receive dte*VT every sub-cycle (FE.cpp:10545-10549): one accumulator, applied
            A.disp[n] = npl[A.np[NP_DSU] * MTP + shs + j] + K.dte * un;         // to both after the loop (k_tauw_owmove)
            A.disp[n + nn] = npl[A.np[NP_DSV] * MTP + sh_nv + j] + K.dte * vn;
        }
    }
    // ghost nodes: moved with the velocity their owner pushed at the end of the previous sub-cycle
    if (A.lag_ghost_move) {
        for (int j = gtid; j < td.n_ghost; j += gstride) {
            int const n = td.ghost_begin + j;
            double u, v;
            mbx_velocity(A.X, n, K.ndof, nn, A.VTc, u, v);
            A.disp[n] += K.dte * u;  A.disp[n + nn] += K.dte * v;
        }
    }
    // every read of this stage is done: one arrival per consumer warp lets the producer refill it
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty + s);
    }   // tile loop

    if (A.fuse_halo) {
        // every sent value of this CTA is on its way: count the CTA in; the last one publishes the epoch in every
        // holder's flag slot (release, system scope) and waits for the owners of this rank's ghosts (bounded spin)
        volatile int& s_last = *(volatile int*)(sm_all + 120);
        __threadfence_system();
        p2_sync();
        if (gtid == 0) s_last = (atomicAdd(A.done_ctr, 1u) == gridDim.x - 1);
        p2_sync();
        if (s_last) {
            __threadfence_system();
            unsigned long long const epoch = *((volatile unsigned long long*)A.epoch_ctr) + 1ULL;
            flag_signal_wait(A.H, gtid, true, epoch, epoch, A.my_flags, 40000000LL, A.halo_err);
            p2_sync();
            if (gtid == 0) { *A.epoch_ctr = epoch; *A.done_ctr = 0u; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Direct path for meshes whose working set is L2-resident (a few 1e5 elements per GPU): one thread per element,
// then one thread per node.  The staging machinery of the tile kernel costs more latency per sub-cycle than it
// saves in traffic when nothing comes from HBM anyway.  Same arithmetic, same summation order (the node kernel
// walks the ascending-element ELL table and subtracts the staged contributions starting from grad_ssh).
// ---------------------------------------------------------------------------------------------------
#ifndef NSX_DIRECT_TPB
#define NSX_DIRECT_TPB 256
#endif
#ifndef NSX_DIRECT_MINB
#define NSX_DIRECT_MINB 1
#endif
constexpr int DIRECT_TPB = NSX_DIRECT_TPB;

// loads of data another SM may have written earlier in the SAME kernel (persistent variant) go to L2 (.cg)
template <int CG> __device__ __forceinline__ double ldv(const double* p) { return CG ? __ldcg(p) : *p; }

struct DirectArgs {
    const int* en0; const int* en1; const int* en2;
    const double* shape; const double* ec;
    double* s0; double* s1; double* s2; double* dm;         // updated in place (persistent) or in -> out
    const double* s0i; const double* s1i; const double* s2i; const double* di;
    double* contrib; const uint8_t* elem_nowrite;
    const uint8_t* nflags; const int* n2e; const int* n2e_deg;
    const double* grad_ssh; const double* node_mass; const double* rlmass; const double* cbu; const double* fcor;
    const double* tau_a; const double* tau_wi; const double* ocean; const double* VTM;
    double* disp;
    MbExchange X;
};

template <int BBM, int CG>
__device__ __forceinline__ void direct_element(KParams const& K, DirectArgs const& A, int e, const double* VT)
{
    int const ne = K.ne, nn = K.nn;
    size_t const NE = (size_t)ne;
    // elements whose writer tile is a boundary tile are written by the tile kernel of the boundary launch; they are
    // still evaluated here because interior nodes next to them need their contributions
    bool const nowrite = A.elem_nowrite && A.elem_nowrite[e];
    double const c0 = A.ec[e];
    double const dx0 = A.shape[e], dx1 = A.shape[NE + e], dx2 = A.shape[2 * NE + e];
    double const dy0 = A.shape[3 * NE + e], dy1 = A.shape[4 * NE + e], dy2 = A.shape[5 * NE + e];
    double s0, s1, s2, vol;
    if (BBM) {
        double const expC = c0;
        vol = A.ec[5 * NE + e];
        double d;
        if (expC == 0.) {                   // conc <= 0.1 : no ice (FE.cpp:4151-4159)
            s0 = s1 = s2 = 0.;
            d = 0.;
        } else {
            int const a = A.en0[e], b = A.en1[e], c = A.en2[e];
            double ua, va, ub, vb, uc, vc;
            mbx_velocity(A.X, a, K.ndof, nn, VT, ua, va);
            mbx_velocity(A.X, b, K.ndof, nn, VT, ub, vb);
            mbx_velocity(A.X, c, K.ndof, nn, VT, uc, vc);
            double e0 = dx0 * ua; e0 += dx1 * ub; e0 += dx2 * uc;
            double e1 = dy0 * va; e1 += dy1 * vb; e1 += dy2 * vc;
            double e2 = dy0 * ua; e2 += dx0 * va; e2 += dy1 * ub; e2 += dx1 * vb; e2 += dy2 * uc; e2 += dx2 * vc;
            s0 = A.s0i[e]; s1 = A.s1i[e]; s2 = A.s2i[e]; d = A.di[e];
            double const dt = K.dte;
            double sigma_n = (s0 + s1) * 0.5;
            double const omd = 1. - d;
            double const time_viscous = K.lambda0 * pow_relax(omd * expC, K);
            double tildeP = 0.;
            if (sigma_n < 0.) tildeP = fmin(1., fast_div(-A.ec[NE + e], sigma_n));
            double const mult = fmin(1. - 1e-12, fast_div(time_viscous, time_viscous + dt * (1. - tildeP)));
            double const elasticity = K.young * omd * expC;
            double const dtE = dt * elasticity;
            s0 += dtE * K.D00 * e0;  s0 += dtE * K.D01 * e1;  s0 *= mult;
            s1 += dtE * K.D01 * e0;  s1 += dtE * K.D00 * e1;  s1 *= mult;
            s2 += dtE * K.D22 * e2;                           s2 *= mult;
            double const sigma_s = fast_hypot((s0 - s1) * 0.5, s2);
            sigma_n = (s0 + s1) * 0.5;
            double dcrit;
            if (sigma_n < -K.compr_strength) dcrit = fast_div(-K.compr_strength, sigma_n);
            else dcrit = fast_div(A.ec[2 * NE + e], sigma_s + K.tan_phi * sigma_n);
            if ((0. < dcrit) && (dcrit < 1.)) {
                double const rtd = fast_sqrt(elasticity) * A.ec[3 * NE + e];
                double const f = (1. - dcrit) * dt * rtd;
                d += omd * f;
                s0 -= s0 * f;  s1 -= s1 * f;  s2 -= s2 * f;
            }
            d = fmax(0., d - A.ec[4 * NE + e]);
        }
        if (!nowrite) A.dm[e] = d;
    } else {
        double const Pp = c0;
        vol = A.ec[NE + e];
        if (Pp < 0.) {                      // thick == 0 (FE.cpp:10656-10662)
            s0 = s1 = s2 = 0.;
        } else {
            int const a = A.en0[e], b = A.en1[e], c = A.en2[e];
            double ua, va, ub, vb, uc, vc;
            mbx_velocity(A.X, a, K.ndof, nn, VT, ua, va);
            mbx_velocity(A.X, b, K.ndof, nn, VT, ub, vb);
            mbx_velocity(A.X, c, K.ndof, nn, VT, uc, vc);
            double eps11 = dx0 * ua; eps11 += dx1 * ub; eps11 += dx2 * uc;
            double eps22 = dy0 * va; eps22 += dy1 * vb; eps22 += dy2 * vc;
            double eps12 = 0.5 * (dx0 * va + dy0 * ua); eps12 += 0.5 * (dx1 * vb + dy1 * ub); eps12 += 0.5 * (dx2 * vc + dy2 * uc);
            double const eps1 = eps11 + eps22, eps2 = eps11 - eps22;
            double const delta = fast_sqrt(eps1 * eps1 + (eps2 * eps2 + 4 * eps12 * eps12) * K.re2);
            double const zeta = fast_div(Pp, delta + K.evp_dmin);
            s0 = A.s0i[e]; s1 = A.s1i[e]; s2 = A.s2i[e];
            double sigma1 = s0 + s1, sigma2 = s0 - s1;
            sigma1 += K.ralpha1 * (zeta * (eps1 - delta) - sigma1);
            sigma2 += K.ralpha2 * (zeta * eps2 * K.re2 - sigma2);
            s2 += K.ralpha2 * (zeta * eps12 * K.re2 - s2);
            s0 = 0.5 * (sigma1 + sigma2);
            s1 = 0.5 * (sigma1 - sigma2);
        }
    }
    if (!nowrite) { A.s0[e] = s0; A.s1[e] = s1; A.s2[e] = s2; }
    // nodal contributions V*(sigma . grad N_i)  (FE.cpp:10464-10465)
    A.contrib[0 * NE + e] = vol * (s0 * dx0 + s2 * dy0);
    A.contrib[1 * NE + e] = vol * (s0 * dx1 + s2 * dy1);
    A.contrib[2 * NE + e] = vol * (s0 * dx2 + s2 * dy2);
    A.contrib[3 * NE + e] = vol * (s2 * dx0 + s1 * dy0);
    A.contrib[4 * NE + e] = vol * (s2 * dx1 + s1 * dy1);
    A.contrib[5 * NE + e] = vol * (s2 * dx2 + s1 * dy2);
}

template <int CG>
__device__ __forceinline__ void direct_node(KParams const& K, DirectArgs const& A, int n, int move_mesh, int lag_ghost_move,
                                            int skip_flag_mask, const double* VTc, double* VTn)
{
    int const nn = K.nn, ne = K.ne;
    uint8_t const fl = A.nflags[n];
    if (fl & skip_flag_mask) return;            // nodes of boundary tiles (and ghosts) belong to the boundary launch
    if (fl & NF_GHOST) {
        if (lag_ghost_move) {
            double u, v;
            mbx_velocity(A.X, n, K.ndof, nn, VTc, u, v);
            A.disp[n] += K.dte * u;  A.disp[n + nn] += K.dte * v;
        }
        return;
    }
    double const uice = ldv<CG>(VTc + n), vice = ldv<CG>(VTc + n + nn);
    double un = uice, vn = vice;
    double const nm = A.node_mass[n];
    if (!(fl & NF_DIRICHLET) && nm != 0.) {
        double gu = A.grad_ssh[n], gv = A.grad_ssh[n + nn];
        int const deg = A.n2e_deg[n];
        for (int k = 0; k < deg; ++k) {             // ascending reference element order (FE.cpp:10445-10467)
            int const s = A.n2e[(size_t)k * nn + n];
            gu -= ldv<CG>(A.contrib + s);
            gv -= ldv<CG>(A.contrib + s + 3 * (size_t)ne);
        }
        double dtep = K.dte, delu = 0., delv = 0.;
        if (K.dynamics_type == NSX_DYN_MEVP) {
            delu = (A.VTM[n] - uice) * K.mevp_rb;
            delv = (A.VTM[n + nn] - vice) * K.mevp_rb;
            dtep = K.dte_mevp;
        }
        double const dte_over_mass = fast_div(dtep, fmax(K.min_m, nm));
        double const ou = A.ocean[n], ov = A.ocean[n + nn];
        double const c_prime = K.rhow_cdw * fast_hypot(ou - uice, ov - vice);
        double const tau_b = A.cbu[n] * fast_div(1., fast_hypot(uice, vice) + K.u0);
        double const sin_s = (fl & NF_LATNEG) ? -K.sin_ota_abs : K.sin_ota_abs;   // std::copysign(sin, lat[i])
        double const alpha = 1. + dte_over_mass * (c_prime * K.cos_ota + tau_b);
        double const beta = dtep * A.fcor[n] + dte_over_mass * c_prime * sin_s;
        double const rdenom = fast_div(1., alpha * alpha + beta * beta);
        double tau_x = A.tau_a[n], tau_y = A.tau_a[n + nn];
        if (A.tau_wi) { tau_x = tau_x + A.tau_wi[n]; tau_y = tau_y + A.tau_wi[n + nn]; }
        tau_x = tau_x + c_prime * (ou * K.cos_ota - ov * sin_s);
        tau_y = tau_y + c_prime * (ov * K.cos_ota + ou * sin_s);
        double const rl = A.rlmass[n];
        double const grad_x = gu * rl, grad_y = gv * rl;
        un = alpha * uice + beta * vice + dte_over_mass * (alpha * (grad_x + tau_x) + beta * (grad_y + tau_y)) + alpha * delu + beta * delv;
        un *= rdenom;
        vn = alpha * vice - beta * uice + dte_over_mass * (alpha * (grad_y + tau_y) - beta * (grad_x + tau_x)) + alpha * delv - beta * delu;
        vn *= rdenom;
    }
    VTn[n] = un;
    VTn[n + nn] = vn;
    if (A.X.on && (fl & NF_BTILE)) mbx_push(A.X, n, un, vn);      // NF_BTILE: node of a tile that owns sent nodes
    if (move_mesh) { A.disp[n] += K.dte * un;  A.disp[n + nn] += K.dte * vn; }
}

template <int BBM>
__global__ void __launch_bounds__(DIRECT_TPB, NSX_DIRECT_MINB)
k_element_direct(KParams K, DirectArgs A, const double* __restrict__ VT)
{
    int const e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.ne) return;
    direct_element<BBM, 0>(K, A, e, VT);
}

__global__ void __launch_bounds__(DIRECT_TPB, NSX_DIRECT_MINB)
k_node_direct(KParams K, DirectArgs A, int move_mesh, int lag_ghost_move, int skip_flag_mask,
              const double* __restrict__ VTc, double* __restrict__ VTn)
{
    int const n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= K.nn) return;
    direct_node<0>(K, A, n, move_mesh, lag_ghost_move, skip_flag_mask, VTc, VTn);
}

// ---------------------------------------------------------------------------------------------------
// State-resident persistent solver for meshes whose sub-cycle state fits the shared memory and registers of the GPU
// (<= ~2.2e5 elements per B200: the headline 10 km mesh, the weak-scaling points, the 3 km mesh on 8 GPUs).
//
// ONE launch runs the whole sub-cycle loop AND the 50 open-water smoother sweeps of a model step.  One CTA per SM owns
// one large tile (ndof / #SMs owned nodes, ~1350 elements + ~10 % redundant halo slots): shape coefficients, rheology
// constants and the stress of every slot stay in shared memory (120 B per slot), damage and the slot connectivity in
// registers of the thread that owns the slot, the velocity of the tile's local nodes in shared memory, UM / UT and the
// incidence list of the owned node in registers.  HBM is not touched inside the loop.
//
// Synchronisation is point to point and carried BY THE DATA (the "LL" scheme of collective libraries): no grid barrier,
// no fence, no separate flag.  Every value another tile or another GPU reads travels as one 16-byte store
// {value, tag} into a MAILBOX slot, tag = exchange epoch; the reader polls the slot (one 16-byte load) until the tag
// is the epoch it expects, so a value is usable the moment it lands.  A tile stores the velocities of its EXPORT nodes
// (owned nodes that another tile or rank reads) into its rank's mailbox, and export nodes on a send list additionally
// straight into the holder's mailbox over NVLink -- this is FiniteElement::updateGhosts (FE.cpp:13963-13996) per
// node; a tile waits only for the ~10 % of its nodes that are halo nodes, each thread for its own.  The mailbox is
// double-buffered by exchange parity.  16-byte aligned vector stores / loads are single transactions on this
// hardware (the property NCCL's LL128 protocol relies on); tags grow monotonically across launches, so no reset.
// The per-sub-cycle schedule hides that latency behind the interior work (plan order: export nodes first, "early"
// slots -- those touching an export or halo node -- first):
//     1. stress of the LATE slots (interior: no halo node, no export node)            needs nothing from outside
//     2. wait for the neighbours' flags of the previous sub-cycle, refresh the halo velocities, move the ghosts
//     3. stress of the EARLY slots
//     4. nodal solve of the EXPORT nodes -> VT buffer (+ NVLink pushes) -> publish
//     5. nodal solve of the other owned nodes (shared memory only)
// Two VT buffers are enough: a neighbour can publish sub-cycle s+1 only after it has read my sub-cycle s, which I
// publish after my last read of sub-cycle s-1 (the neighbour relation is symmetric, between tiles and between ranks).
// Same arithmetic and summation order as the other paths; the nodal contributions are recomputed from the resident
// stress instead of being staged, so results may differ from them in the last bit (FMA contraction).
// All spins are bounded and report through the halo error word.
// ---------------------------------------------------------------------------------------------------
#ifndef NSX_RES_TPB
#define NSX_RES_TPB 768
#endif
#ifndef NSX_RES_CTAS
#define NSX_RES_CTAS 1
#endif
constexpr int RES_TPB = NSX_RES_TPB;            // threads per tile: one owned node per thread, up to RES_SPT slots per thread
constexpr int RES_CTAS = NSX_RES_CTAS;          // tiles (CTAs) per SM: independent tiles fill each other's barrier / latency stalls
constexpr int RES_SPT = 3;                      // slots per thread (static unroll): tiles of up to 3 * RES_TPB slots
constexpr int RES_MAX_LINKS = MB_MAX_PEERS;     // neighbour ranks of one rank

struct ResPeers {
    int n_send;                                 // neighbour ranks this rank sends to (send slot = push_ent.x)
    MbEntry* send_mb[RES_MAX_LINKS];            // the holder's mailbox (parity 0; parity 1 follows send_nmb entries later)
    int send_nmb[RES_MAX_LINKS];
};

struct ResidentArgs {
    const TileDesc* tiles; const ResTile* rtiles;
    const int* halo_nodes; const int* halo_slot; const uint8_t* halo_move; const int* halo_elems; const unsigned long long* slot_conn;
    const double* slot_shape; const double* slot_ec; int nslots; const uint16_t* inc;
    const uint16_t* n2n_loc; const uint8_t* n2n_deg;
    double* s0; double* s1; double* s2; double* dm;               // updated in place at the end of the loop
    const uint8_t* nflags; const double* grad_ssh; const double* node_mass; const double* rlmass; const double* cbu;
    const double* fcor; const double* tau_a; const double* tau_wi; const double* ocean; const double* VTM;
    double* VT0; double* VT1; int cur; double* UM; double* UT;
    int move_mesh, nsub, nsweeps;
    const int* ow_count;
    MbEntry* mb; int n_mb;                                        // this rank's mailbox: [2][n_mb], export nodes then ghost nodes
    int has_peers;                                                // neighbour ranks exist: the smoother sweeps are always exchanged
    const int* push_ptr; const int2* push_ent;                    // owned node -> (send slot, slot in the holder's mailbox)
    const unsigned long long* epoch_ctr; int* err;
    unsigned long long* tstamp;                                   // [3] %globaltimer: launch start, end of the sub-cycle loop, end (max over tiles)
    int MS, MLN;                                                  // shared-memory strides: slots per plane, local nodes
    ResPeers P;
};

// stress (and damage) update of one resident slot; returns the new damage
template <int BBM>
__device__ __forceinline__ double res_slot_update(KParams const& K, int k, unsigned long long pc, double d, int MS,
                                                  const double* shp, const double* ecp, double* sgp, const double* su, const double* sv)
{
    int const la = (int)(pc & 0xFFFF), lb = (int)((pc >> 16) & 0xFFFF), lc = (int)((pc >> 32) & 0xFFFF);
    double const dx0 = shp[k], dx1 = shp[MS + k], dx2 = shp[2 * MS + k];
    double const dy0 = shp[3 * MS + k], dy1 = shp[4 * MS + k], dy2 = shp[5 * MS + k];
    double const c0 = ecp[k];
    double s0, s1, s2;
    if (BBM) {
        double const expC = c0;
        if (expC == 0.) {                   // conc <= 0.1 : no ice (FE.cpp:4151-4159)
            s0 = s1 = s2 = 0.;
            d = 0.;
        } else {
            double const ua = su[la], va = sv[la], ub = su[lb], vb = sv[lb], uc = su[lc], vc = sv[lc];
            double e0 = dx0 * ua; e0 += dx1 * ub; e0 += dx2 * uc;
            double e1 = dy0 * va; e1 += dy1 * vb; e1 += dy2 * vc;
            double e2 = dy0 * ua; e2 += dx0 * va; e2 += dy1 * ub; e2 += dx1 * vb; e2 += dy2 * uc; e2 += dx2 * vc;
            s0 = sgp[k]; s1 = sgp[MS + k]; s2 = sgp[2 * MS + k];
            double const dt = K.dte;
            double sigma_n = (s0 + s1) * 0.5;
            double const omd = 1. - d;
            double const time_viscous = K.lambda0 * pow_relax(omd * expC, K);
            double tildeP = 0.;
            if (sigma_n < 0.) tildeP = fmin(1., fast_div(-ecp[MS + k], sigma_n));
            double const mult = fmin(1. - 1e-12, fast_div(time_viscous, time_viscous + dt * (1. - tildeP)));
            double const elasticity = K.young * omd * expC;
            double const dtE = dt * elasticity;
            s0 += dtE * K.D00 * e0;  s0 += dtE * K.D01 * e1;  s0 *= mult;
            s1 += dtE * K.D01 * e0;  s1 += dtE * K.D00 * e1;  s1 *= mult;
            s2 += dtE * K.D22 * e2;                           s2 *= mult;
            double const sigma_s = fast_hypot((s0 - s1) * 0.5, s2);
            sigma_n = (s0 + s1) * 0.5;
            double dcrit;
            if (sigma_n < -K.compr_strength) dcrit = fast_div(-K.compr_strength, sigma_n);
            else dcrit = fast_div(ecp[2 * MS + k], sigma_s + K.tan_phi * sigma_n);
            if ((0. < dcrit) && (dcrit < 1.)) {
                double const rtd = fast_sqrt(elasticity) * ecp[3 * MS + k];
                double const f = (1. - dcrit) * dt * rtd;
                d += omd * f;
                s0 -= s0 * f;  s1 -= s1 * f;  s2 -= s2 * f;
            }
            d = fmax(0., d - ecp[4 * MS + k]);
        }
    } else {
        double const Pp = c0;
        if (Pp < 0.) {                      // thick == 0 (FE.cpp:10656-10662)
            s0 = s1 = s2 = 0.;
        } else {
            double const ua = su[la], va = sv[la], ub = su[lb], vb = sv[lb], uc = su[lc], vc = sv[lc];
            double eps11 = dx0 * ua; eps11 += dx1 * ub; eps11 += dx2 * uc;
            double eps22 = dy0 * va; eps22 += dy1 * vb; eps22 += dy2 * vc;
            double eps12 = 0.5 * (dx0 * va + dy0 * ua); eps12 += 0.5 * (dx1 * vb + dy1 * ub); eps12 += 0.5 * (dx2 * vc + dy2 * uc);
            double const eps1 = eps11 + eps22, eps2 = eps11 - eps22;
            double const delta = fast_sqrt(eps1 * eps1 + (eps2 * eps2 + 4 * eps12 * eps12) * K.re2);
            double const zeta = fast_div(Pp, delta + K.evp_dmin);
            s0 = sgp[k]; s1 = sgp[MS + k]; s2 = sgp[2 * MS + k];
            double sigma1 = s0 + s1, sigma2 = s0 - s1;
            sigma1 += K.ralpha1 * (zeta * (eps1 - delta) - sigma1);
            sigma2 += K.ralpha2 * (zeta * eps2 * K.re2 - sigma2);
            s2 += K.ralpha2 * (zeta * eps12 * K.re2 - s2);
            s0 = 0.5 * (sigma1 + sigma2);
            s1 = 0.5 * (sigma1 - sigma2);
        }
    }
    sgp[k] = s0; sgp[MS + k] = s1; sgp[2 * MS + k] = s2;
    return d;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int BBM>
__global__ void __launch_bounds__(RES_TPB, RES_CTAS)
k_resident(KParams K, ResidentArgs A)
{
    extern __shared__ __align__(16) unsigned char sm_res[];
    int const nn = K.nn, tid = threadIdx.x, MS = A.MS, MLN = A.MLN;
    constexpr int NEC = BBM ? 6 : 2;
    double* const shp = (double*)sm_res;                          // [6][MS]
    double* const ecp = shp + 6 * (size_t)MS;                     // [NEC][MS]
    double* const sgp = ecp + NEC * (size_t)MS;                   // [3][MS]
    unsigned long long* const cnp = (unsigned long long*)(sgp + 3 * (size_t)MS);   // [MS] slot connectivity (3 x u16 local ids)
    double* const su = (double*)(cnp + MS);                       // [MLN]
    double* const sv = su + MLN;                                  // [MLN]
    // tile descriptors live in shared memory (broadcast reads): registers are the scarce resource of this kernel
    __shared__ TileDesc td;
    __shared__ ResTile rt;
    if (tid == 0) { td = A.tiles[blockIdx.x]; rt = A.rtiles[blockIdx.x]; }
    __syncthreads();
    int const nsl = td.n_own_slots + td.n_halo_slots;
    size_t const NS = (size_t)A.nslots;
    ResPeers const& P = A.P;
    unsigned long long const epoch0 = *A.epoch_ctr;               // advanced by the host-side sequence after every launch

    if (tid == 0 && blockIdx.x == 0) A.tstamp[0] = globaltimer_ns();
    // Which slots a thread updates.  Round 0 runs BEFORE the halo of the previous sub-cycle is awaited: one late slot (no
    // halo node, no export node) per thread, k = n_early_own + tid.  Rounds 1, 2 run after it and take everything else in
    // order -- early own slots, the late slots beyond the first RES_TPB, halo slots.  A balanced tile (<= 2 * RES_TPB slots,
    // at least nsl - RES_TPB late ones) needs one round on either side of the wait, and every warp has work in both.
    int const late_end = min(td.n_own_slots, rt.n_early_own + RES_TPB);
    int const n_after = rt.n_early_own + nsl - late_end;          // <= 2 * RES_TPB, checked by the host
    auto slot_of = [&](int round) -> int {
        if (round == 0) { int const k = rt.n_early_own + tid; return k < late_end ? k : -1; }
        int const idx = tid + (round - 1) * RES_TPB;
        if (idx >= n_after) return -1;
        return idx < rt.n_early_own ? idx : late_end + (idx - rt.n_early_own);
    };
    // ---- load the tile once ----
    double dmg[RES_SPT];
#pragma unroll
    for (int q = 0; q < RES_SPT; ++q) {
        dmg[q] = 0.;
        if (BBM) {
            int const k = slot_of(q);
            if (k >= 0) dmg[q] = A.dm[(k < td.n_own_slots) ? td.elem_begin + k : A.halo_elems[td.halo_elem_off + (k - td.n_own_slots)]];
        }
    }
#pragma unroll
    for (int q = 0; q < RES_SPT; ++q) {
        int const k = tid + q * RES_TPB;
        if (k < nsl) {
            size_t const g = (size_t)td.slot_begin + k;
            cnp[k] = A.slot_conn[g];
            {   // four planes in slot space (see k_prep_elements); the resident copy keeps all six
                double const dx0 = A.slot_shape[0 * NS + g], dx1 = A.slot_shape[1 * NS + g];
                double const dy0 = A.slot_shape[2 * NS + g], dy1 = A.slot_shape[3 * NS + g];
                shp[k] = dx0; shp[MS + k] = dx1; shp[2 * MS + k] = -(dx0 + dx1);
                shp[3 * MS + k] = dy0; shp[4 * MS + k] = dy1; shp[5 * MS + k] = -(dy0 + dy1);
            }
#pragma unroll
            for (int c = 0; c < NEC; ++c) ecp[c * MS + k] = A.slot_ec[c * NS + g];
            int const e = (k < td.n_own_slots) ? td.elem_begin + k : A.halo_elems[td.halo_elem_off + (k - td.n_own_slots)];
            sgp[k] = A.s0[e]; sgp[MS + k] = A.s1[e]; sgp[2 * MS + k] = A.s2[e];
        }
    }
    auto VTb = [&](int parity) -> double* { return parity ? A.VT1 : A.VT0; };
    {
        const double* VTr = VTb(A.cur & 1);
        for (int j = tid; j < td.n_own; j += RES_TPB) { su[j] = VTr[td.node_begin + j]; sv[j] = VTr[td.node_begin + j + nn]; }
        for (int h = tid; h < td.n_halo; h += RES_TPB) {
            int const g = A.halo_nodes[td.halo_off + h];
            su[td.n_own + HALO_GAP + h] = VTr[g]; sv[td.n_own + HALO_GAP + h] = VTr[g + nn];
        }
    }
    bool const has_node = tid < td.n_own;                         // one owned node per thread (checked by the host)
    int const n = td.node_begin + (has_node ? tid : 0);
    uint8_t const fl = has_node ? A.nflags[n] : (uint8_t)NF_DIRICHLET;
    double const nmass = has_node ? A.node_mass[n] : 1.;
    bool const solve_node = has_node && !(fl & NF_DIRICHLET) && nmass != 0.;
    // dte / max(min_m, mass) of the owned node is the same in every sub-cycle (FE.cpp:10483): one division per launch
    double const dte_over_mass_c = fast_div(K.dynamics_type == NSX_DYN_MEVP ? K.dte_mevp : K.dte, fmax(K.min_m, nmass));
    bool const mass_is_zero = nmass == 0.;
    // displacement of the owned node over the launch: M_UM and M_UT receive the same increments dte*VT every sub-cycle
    // (FE.cpp:10545-10549), so one accumulator serves both (added once at the end; M_UM of Neumann nodes is restored)
    double dspu = 0., dspv = 0.;
    // incidence list of the owned node, pre-decoded (slot | vertex << 14), 8 entries in registers
    unsigned int incr[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (has_node) {
        const uint16_t* ip = A.inc + td.inc_off + tid;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c >= td.inc_w) break;
            unsigned const code = ip[(size_t)c * td.n_own];
            unsigned dec = 0xFFFFu;
            if (code != 0xFFFFu) { unsigned const i = code / (unsigned)MS; dec = (code - i * (unsigned)MS) | (i << 14); }
            incr[c >> 1] = (c & 1) ? ((incr[c >> 1] & 0x0000FFFFu) | (dec << 16)) : ((incr[c >> 1] & 0xFFFF0000u) | dec);
        }
    }
    // the first halo node this thread refreshes: its mailbox slot, and (ghost nodes this tile moves, FE.cpp:10539-10553)
    // its node id, -1 otherwise
    int const hslot0 = (tid < td.n_halo) ? A.halo_slot[td.halo_off + tid] : 0;
    int const hmove0 = (tid < td.n_halo && A.halo_move[td.halo_off + tid]) ? A.halo_nodes[td.halo_off + tid] : -1;
    __syncthreads();

    // ---- helpers ----
    auto phase1 = [&](bool after_wait) {
#pragma unroll
        for (int q = 0; q < RES_SPT; ++q) {
            if ((q > 0) != after_wait) continue;
            int const k = slot_of(q);
            if (k < 0) continue;
            NSX_DEV_CHECK((int)(cnp[k] & 0xFFFF) < MLN && (int)((cnp[k] >> 16) & 0xFFFF) < MLN && (int)((cnp[k] >> 32) & 0xFFFF) < MLN && k < MS, A.err, 1);
            dmg[q] = res_slot_update<BBM>(K, k, cnp[k], dmg[q], MS, shp, ecp, sgp, su, sv);
        }
    };
    // nodal solve of this thread's node (FE.cpp:10445-10529): contributions in ascending reference element order
    // the nine step-constant values of the owned node come from L2 every sub-cycle (no register or shared-memory room to
    // keep them): the loads are issued BEFORE the barrier that precedes the solve, so their latency overlaps the wait
    struct NodeConst { double gu, gv, rl, cb, fc, tau_x, tau_y, ou, ov; };
    auto node_consts = [&]() {
        NodeConst c;
        c.gu = c.gv = c.rl = c.cb = c.fc = c.tau_x = c.tau_y = c.ou = c.ov = 0.;
        if (!solve_node) return c;
        c.gu = __ldg(A.grad_ssh + n); c.gv = __ldg(A.grad_ssh + n + nn);
        c.rl = __ldg(A.rlmass + n); c.cb = __ldg(A.cbu + n); c.fc = __ldg(A.fcor + n);
        c.tau_x = __ldg(A.tau_a + n); c.tau_y = __ldg(A.tau_a + n + nn);
        c.ou = __ldg(A.ocean + n); c.ov = __ldg(A.ocean + n + nn);
        if (A.tau_wi) { c.tau_x = c.tau_x + __ldg(A.tau_wi + n); c.tau_y = c.tau_y + __ldg(A.tau_wi + n + nn); }
        return c;
    };
    auto node_solve = [&](NodeConst const& nc, double& un, double& vn) {
        double const uice = su[tid], vice = sv[tid];
        un = uice; vn = vice;
        if (!solve_node) return;
        double gu = nc.gu, gv = nc.gv;
        double const rl = nc.rl, cb = nc.cb, fc = nc.fc;
        double tau_x = nc.tau_x, tau_y = nc.tau_y;
        double const ou = nc.ou, ov = nc.ov;
        bool more = true;
        // keep the packed codes opaque: otherwise the compiler hoists the 8 x 3 decoded shared-memory addresses out of the
        // sub-cycle loop and spills them (registers are the scarce resource here; the decode is two shifts)
        asm volatile("" : "+r"(incr[0]), "+r"(incr[1]), "+r"(incr[2]), "+r"(incr[3]));
        // two incidences at a time: their twelve shared-memory loads and their products are independent, only the two
        // subtractions keep the reference's order (an invalid second entry reads the first one's slot and is dropped)
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            unsigned const code0 = incr[c >> 1] & 0xFFFFu, code1 = incr[c >> 1] >> 16;
            if (code0 == 0xFFFFu) { more = false; break; }
            bool const two = code1 != 0xFFFFu;
            unsigned const cd1 = two ? code1 : code0;
            int const i0 = (int)(code0 >> 14), k0 = (int)(code0 & 0x3FFFu);
            int const i1 = (int)(cd1 >> 14), k1 = (int)(cd1 & 0x3FFFu);
            NSX_DEV_CHECK(i0 < 3 && k0 < nsl && i1 < 3 && k1 < nsl, A.err, 2);
            double const a0 = sgp[k0], a1 = sgp[MS + k0], a2 = sgp[2 * MS + k0];
            double const vol = ecp[(BBM ? 5 : 1) * MS + k0];
            double const dxi = shp[i0 * MS + k0], dyi = shp[(3 + i0) * MS + k0];
            double const b0 = sgp[k1], b1 = sgp[MS + k1], b2 = sgp[2 * MS + k1];
            double const wol = ecp[(BBM ? 5 : 1) * MS + k1];
            double const exi = shp[i1 * MS + k1], eyi = shp[(3 + i1) * MS + k1];
            double const pu0 = a0 * dxi + a2 * dyi, pv0 = a2 * dxi + a1 * dyi;       // sigma . grad N_i, FE.cpp:10464-10465
            double const pu1 = b0 * exi + b2 * eyi, pv1 = b2 * exi + b1 * eyi;
            gu = fma(-vol, pu0, gu);                      // gu -= V * (...): the fused form the single-entry loop compiled to
            gv = fma(-vol, pv0, gv);
            if (!two) { more = false; break; }
            gu = fma(-wol, pu1, gu);
            gv = fma(-wol, pv1, gv);
        }
        if (more && td.inc_w > 8) {                      // nodes with more than 8 incident elements: rest from the table
            const uint16_t* ip = A.inc + td.inc_off + tid;
            for (int c = 8; c < td.inc_w; ++c) {
                unsigned const code = __ldg(ip + (size_t)c * td.n_own);
                if (code == 0xFFFFu) break;
                int const i = (int)(code / (unsigned)MS), k = (int)(code - (unsigned)i * (unsigned)MS);
                double const a0 = sgp[k], a1 = sgp[MS + k], a2 = sgp[2 * MS + k];
                double const vol = ecp[(BBM ? 5 : 1) * MS + k];
                double const dxi = shp[i * MS + k], dyi = shp[(3 + i) * MS + k];
                gu -= vol * (a0 * dxi + a2 * dyi);
                gv -= vol * (a2 * dxi + a1 * dyi);
            }
        }
        double dtep = K.dte, delu = 0., delv = 0.;
        if (K.dynamics_type == NSX_DYN_MEVP) {
            delu = (__ldg(A.VTM + n) - uice) * K.mevp_rb;
            delv = (__ldg(A.VTM + n + nn) - vice) * K.mevp_rb;
            dtep = K.dte_mevp;
        }
        double const dte_over_mass = dte_over_mass_c;
        double const c_prime = K.rhow_cdw * fast_hypot(ou - uice, ov - vice);
        double const tau_b = cb * fast_div(1., fast_hypot(uice, vice) + K.u0);
        double const sin_s = (fl & NF_LATNEG) ? -K.sin_ota_abs : K.sin_ota_abs;
        double const alpha = 1. + dte_over_mass * (c_prime * K.cos_ota + tau_b);
        double const beta = dtep * fc + dte_over_mass * c_prime * sin_s;
        double const rdenom = fast_div(1., alpha * alpha + beta * beta);
        tau_x = tau_x + c_prime * (ou * K.cos_ota - ov * sin_s);
        tau_y = tau_y + c_prime * (ov * K.cos_ota + ou * sin_s);
        double const grad_x = gu * rl, grad_y = gv * rl;
        un = alpha * uice + beta * vice + dte_over_mass * (alpha * (grad_x + tau_x) + beta * (grad_y + tau_y)) + alpha * delu + beta * delv;
        un *= rdenom;
        vn = alpha * vice - beta * uice + dte_over_mass * (alpha * (grad_y + tau_y) - beta * (grad_x + tau_x)) + alpha * delv - beta * delu;
        vn *= rdenom;
    };
    // export node -> this rank's mailbox and, if it is on a send list, the mailbox of every holder (updateGhosts,
    // FE.cpp:13963-13996): value and exchange epoch in one 16-byte store each
    auto publish_node = [&](int parity, int ex, double un, double vn) {
        unsigned long long const tag = epoch0 + (unsigned long long)ex;
        NSX_DEV_CHECK(rt.x_off + tid < A.n_mb && n < K.ndof, A.err, 3);
        mb_store(A.mb + (size_t)parity * A.n_mb + rt.x_off + tid, un, vn, tag);
        if (P.n_send) {
            int const q1 = A.push_ptr[n + 1];
            for (int q = A.push_ptr[n]; q < q1; ++q) {
                int2 const pe = A.push_ent[q];
                NSX_DEV_CHECK(pe.x >= 0 && pe.x < P.n_send && pe.y >= 0 && pe.y < P.send_nmb[pe.x], A.err, 4);
                mb_store(P.send_mb[pe.x] + (size_t)parity * P.send_nmb[pe.x] + pe.y, un, vn, tag);
            }
        }
    };
    // every thread waits for ITS halo node of exchange ex (tile of this GPU or ghost of another one) and refreshes it
    // displacement of the ghost node this thread moves (its first halo entry): accumulated like dspu / dspv, written once
    double gdu = 0., gdv = 0.;
    auto wait_refresh = [&](int ex, int parity, bool move_ghosts, double dt_move) {
        unsigned long long const tag = epoch0 + (unsigned long long)ex;
        const MbEntry* const box = A.mb + (size_t)parity * A.n_mb;
        for (int h = tid; h < td.n_halo; h += RES_TPB) {
            int const slot = (h == tid) ? hslot0 : A.halo_slot[td.halo_off + h];
            NSX_DEV_CHECK(slot >= 0 && slot < A.n_mb && td.n_own + HALO_GAP + h < MLN, A.err, 5);
            double u, v;
            mb_wait(box + slot, tag, u, v, A.err, 1000 + (int)blockIdx.x);
            su[td.n_own + HALO_GAP + h] = u; sv[td.n_own + HALO_GAP + h] = v;
            if (!move_ghosts) continue;                  // ghost nodes move with their owner's fresh velocity
            if (h == tid) { gdu = gdu + dt_move * u;  gdv = gdv + dt_move * v; continue; }
            if (!A.halo_move[td.halo_off + h]) continue; // tiles with more halo nodes than threads: read-modify-write
            int const g = A.halo_nodes[td.halo_off + h];
            A.UT[g] += dt_move * u;  A.UT[g + nn] += dt_move * v;
            if (!(A.nflags[g] & NF_NEUMANN)) { A.UM[g] += dt_move * u;  A.UM[g + nn] += dt_move * v; }
        }
        __syncthreads();
    };

    // ---- the sub-cycle loop (FE.cpp:10423-10554) ----
    int const cur = A.cur & 1;
    for (int s = 0; s < A.nsub; ++s) {
        int const pw = (cur + s + 1) & 1;                         // parity of the buffer this sub-cycle writes
        phase1(false);                                            // 1. one late slot per thread
        if (s > 0) wait_refresh(s, pw ^ 1, A.move_mesh != 0, K.dte);   // 2. (sub-cycle 0 starts from the loaded state)
        else __syncthreads();                                     // (the barrier also orders phase 1 before the node solve)
        phase1(true);                                             // 3. early, remaining late and halo slots
        NodeConst const nc = node_consts();
        __syncthreads();
        double un = 0., vn = 0.;
        if (has_node) node_solve(nc, un, vn);                     // needs every incident stress: after both phase-1 parts
        // 4./5. one node per thread: every solve has read its inputs, nobody else reads su/sv before the barrier below
        if (has_node) { su[tid] = un; sv[tid] = vn; }
        if (tid < rt.n_x) publish_node(pw, s + 1, un, vn);        // export nodes -> mailboxes (local + NVLink), usable on arrival
        if (has_node && A.move_mesh) { dspu = dspu + K.dte * un;  dspv = dspv + K.dte * vn; }
        __syncthreads();                                          // su / sv complete before the next phase 1 reads them
    }
    int ex = A.nsub;                                              // exchanges done so far
    if (ex > 0) wait_refresh(ex, (cur + ex) & 1, A.move_mesh != 0, K.dte);
    if (tid == 0) atomicMax(A.tstamp + 1, globaltimer_ns());
    if (K.dynamics_type == NSX_DYN_MEVP && ex > 0) {
        // mEVP: ONE mesh move with the full time step after the loop (FE.cpp:10559-10573); ghosts alike
        if (has_node) { dspu = K.dtime_step * su[tid];  dspv = K.dtime_step * sv[tid]; }
        for (int h = tid; h < td.n_halo; h += RES_TPB) {
            double const u = su[td.n_own + HALO_GAP + h], v = sv[td.n_own + HALO_GAP + h];
            if (h == tid) { gdu = K.dtime_step * u;  gdv = K.dtime_step * v; continue; }
            if (!A.halo_move[td.halo_off + h]) continue;
            int const g = A.halo_nodes[td.halo_off + h];
            A.UT[g] += K.dtime_step * u;  A.UT[g + nn] += K.dtime_step * v;
            if (!(A.nflags[g] & NF_NEUMANN)) { A.UM[g] += K.dtime_step * u;  A.UM[g + nn] += K.dtime_step * v; }
        }
    }

    // ---- open-water smoother: 50 Jacobi sweeps over the ice-free nodes (FE.cpp:10578-10611), same exchange per sweep.
    // A rank without neighbours and without ice-free nodes skips it; with neighbour ranks every sweep is exchanged.
    int const nsweeps = (A.nsweeps > 0 && (A.has_peers || *A.ow_count > 0)) ? A.nsweeps : 0;
    bool const is_ow = has_node && !(fl & NF_DIRICHLET) && mass_is_zero;
    int const deg = is_ow ? (int)A.n2n_deg[n] : 0;
    // the first eight neighbour ids of an open-water node move into the registers that held the incidence codes: the
    // sweeps then read no global memory (an L2 round trip per sweep otherwise)
    if (nsweeps > 0 && is_ow) {
        const uint16_t* q = A.n2n_loc + rt.n2n_off + tid;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned const l = j < deg ? (unsigned)__ldg(q + (size_t)j * td.n_own) : 0xFFFFu;
            incr[j >> 1] = (j & 1) ? ((incr[j >> 1] & 0x0000FFFFu) | (l << 16)) : ((incr[j >> 1] & 0xFFFF0000u) | l);
        }
    }
    for (int it = 0; it < nsweeps; ++it) {
        int const pw = (cur + ex + 1) & 1;
        double nu = 0., nv = 0.;
        if (is_ow) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                unsigned const l = (incr[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                if (j >= deg) break;
                nu += su[l]; nv += sv[l];
            }
            if (deg > 8) {
                const uint16_t* q = A.n2n_loc + rt.n2n_off + tid;
                for (int j = 8; j < deg; ++j) {
                    int const l = (int)__ldg(q + (size_t)j * td.n_own);
                    nu += su[l]; nv += sv[l];
                }
            }
            nu = nu / deg; nv = nv / deg;
        }
        __syncthreads();                                          // Jacobi: every read of the old values is done
        if (is_ow) { su[tid] = nu; sv[tid] = nv; }
        ++ex;
        if (tid < rt.n_x) publish_node(pw, ex, su[tid], sv[tid]);
        wait_refresh(ex, pw, false, 0.);
    }

    // ---- write the resident state back ----
    if (tid == 0) atomicMax(A.tstamp + 2, globaltimer_ns());
    {
        // always the buffer of parity cur + nsub + nsweeps: the host cannot know whether the sweeps were skipped (only a
        // rank without neighbours, i.e. without ghosts to keep current, ever skips them)
        double* VTf = VTb((cur + A.nsub + A.nsweeps) & 1);
        if (has_node) {
            VTf[n] = su[tid]; VTf[n + nn] = sv[tid];
            A.UT[n] += dspu;  A.UT[n + nn] += dspv;
            if (!(fl & NF_NEUMANN)) { A.UM[n] += dspu;  A.UM[n + nn] += dspv; }
        }
        // ghost nodes: their last received velocity goes to the VT buffer the host-side sequence continues with
        for (int h = tid; h < td.n_halo; h += RES_TPB) {
            if (!A.halo_move[td.halo_off + h]) continue;
            int const g = A.halo_nodes[td.halo_off + h];
            VTf[g] = su[td.n_own + HALO_GAP + h]; VTf[g + nn] = sv[td.n_own + HALO_GAP + h];
            if (h == tid) {
                A.UT[g] += gdu;  A.UT[g + nn] += gdv;
                if (!(A.nflags[g] & NF_NEUMANN)) { A.UM[g] += gdu;  A.UM[g + nn] += gdv; }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < RES_SPT; ++q) {
        int const k = tid + q * RES_TPB;
        if (k < td.n_own_slots) {
            int const e = td.elem_begin + k;
            A.s0[e] = sgp[k]; A.s1[e] = sgp[MS + k]; A.s2[e] = sgp[2 * MS + k];
        }
    }
    if (BBM) {
#pragma unroll
        for (int q = 0; q < RES_SPT; ++q) {
            int const k = slot_of(q);
            if (k >= 0 && k < td.n_own_slots) A.dm[td.elem_begin + k] = dmg[q];
        }
    }
}

// mesh move over an explicit node range with an explicit time increment:
//   mEVP: once after the loop with dtime_step (FE.cpp:10559-10573); ghosts: final lagged move.
__global__ void __launch_bounds__(TPB)
k_move_mesh(int nn, int n_begin, int n_end, double dt, const double* __restrict__ VT, double* __restrict__ disp)
{
    int const n = n_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_end) return;
    disp[n] += dt * VT[n];  disp[n + nn] += dt * VT[n + nn];
}

// ---------------------------------------------------------------------------------------------------
// open-water smoother, one Jacobi sweep over the compacted list (FE.cpp:10580-10608)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_ow_sweep(int nn, const int* __restrict__ ow_list, const int* __restrict__ ow_count,
           const int* __restrict__ n2n, const int* __restrict__ n2n_deg,
           const double* __restrict__ VTin, double* __restrict__ VTout)
{
    int const cnt = *ow_count;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += gridDim.x * blockDim.x) {
        int const n = ow_list[t];
        int const deg = n2n_deg[n];
        double su = 0., sv = 0.;
        for (int j = 0; j < deg; ++j) {
            int const q = n2n[(size_t)j * nn + n];
            su += VTin[q];
            sv += VTin[q + nn];
        }
        VTout[n] = su / deg;
        VTout[n + nn] = sv / deg;
    }
}

// All 50 Jacobi sweeps in ONE launch for a rank without neighbours: a software grid barrier (the grid is at most one
// CTA per SM, so every CTA is resident) separates the sweeps.  Exits at once when there is no open-water node.
__global__ void __launch_bounds__(TPB)
k_ow_smooth_all(int nn, int nsweeps, const int* __restrict__ ow_list, const int* __restrict__ ow_count,
                const int* __restrict__ n2n, const int* __restrict__ n2n_deg,
                double* VTa, double* VTb, unsigned int* bar, int* err)
{
    int const cnt = *ow_count;
    if (cnt == 0) return;
    // only as many CTAs as the list needs take part (a barrier over few CTAs is cheaper)
    int const nact = min((int)gridDim.x, (cnt + (int)blockDim.x - 1) / (int)blockDim.x);
    if ((int)blockIdx.x >= nact) return;
    for (int it = 0; it < nsweeps; ++it) {
        const double* VTin = (it & 1) ? VTb : VTa;
        double* VTout = (it & 1) ? VTa : VTb;
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += nact * blockDim.x) {
            int const n = ow_list[t];
            int const deg = n2n_deg[n];
            double su = 0., sv = 0.;
            for (int j = 0; j < deg; ++j) {
                int const q = n2n[(size_t)j * nn + n];
                su += __ldcg(VTin + q);              // L2: written by other SMs in the previous sweep
                sv += __ldcg(VTin + q + nn);
            }
            VTout[n] = su / deg;
            VTout[n + nn] = sv / deg;
        }
        // grid barrier over the nact participating CTAs: release my sweep, acquire everybody else's.  The spin is
        // bounded (a grid that is not co-resident must not hang the device) and a timeout sets the error word that
        // nsx_download / nsx_check report.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            unsigned int const target = (unsigned int)(it + 1) * (unsigned int)nact;
            unsigned int v;
            long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
                if (v >= target) break;
                if (++spins > (1LL << 24)) { atomicExch(err, 2000); break; }
            }
            __threadfence();
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// ice-ocean stress diagnostic + open-water mesh move  (FE.cpp:10615-10640)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_tauw_owmove(KParams K, const uint8_t* __restrict__ nflags, const double* __restrict__ node_mass,
              const double* __restrict__ VT, const double* __restrict__ VTM, const double* __restrict__ ocean,
              double* __restrict__ tau_w, double* __restrict__ UM, double* __restrict__ UT, double* __restrict__ disp,
              unsigned long long* epoch_ctr, unsigned long long epoch_bump)
{
    int const n = blockIdx.x * blockDim.x + threadIdx.x;
    int const nn = K.nn;
    if (n == 0 && epoch_ctr) *epoch_ctr += epoch_bump;       // exchanges done by the resident launch that just finished
    if (n >= nn) return;
    uint8_t const fl0 = nflags[n];
    if (disp) {     // displacement accumulated by the sub-cycle loop of the tile / direct paths: M_UT and M_UM (Neumann nodes
        double const du = disp[n], dv = disp[n + nn];          // keep their M_UM, FE.cpp:10551-10552) receive it once
        UT[n] += du;  UT[n + nn] += dv;
        if (!(fl0 & NF_NEUMANN)) { UM[n] += du;  UM[n + nn] += dv; }
        disp[n] = 0.;  disp[n + nn] = 0.;
    }
    double const u = VT[n], v = VT[n + nn];
    double const uice = 0.5 * (u + VTM[n]);
    double const vice = 0.5 * (v + VTM[n + nn]);
    double const ou = ocean[n], ov = ocean[n + nn];
    double const c_prime = K.rhow_cdw * hypot(ou - uice, ov - vice);
    tau_w[n] = c_prime * (uice - ou);
    tau_w[n + nn] = c_prime * (vice - ov);
    uint8_t const fl = nflags[n];
    if ((fl & NF_DIRICHLET) || node_mass[n] != 0.) return;
    UT[n] += K.dtime_step * u;  UT[n + nn] += K.dtime_step * v;
    if (!(fl & NF_NEUMANN)) { UM[n] += K.dtime_step * u;  UM[n + nn] += K.dtime_step * v; }
}

// ---------------------------------------------------------------------------------------------------
// update()  (FE.cpp:3946-4131): Lagrangian area change, ridging, mechanical redistribution, clamps
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_update(KParams K, const uint8_t* __restrict__ nflags,
         const int* __restrict__ en0, const int* __restrict__ en1, const int* __restrict__ en2,
         const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ UM,
         double* __restrict__ surface, double* __restrict__ conc, double* __restrict__ thick, double* __restrict__ snow,
         double* __restrict__ thick_myi, double* __restrict__ conc_myi, double* __restrict__ ridge_ratio,
         double* __restrict__ conc_y, double* __restrict__ h_y, double* __restrict__ hs_y,
         double* __restrict__ sig0, double* __restrict__ sig1, double* __restrict__ sig2,
         double* __restrict__ del_ci_ridge_myi)
{
    int const e = blockIdx.x * blockDim.x + threadIdx.x;
    int const nn = K.nn;
    if (e >= K.ne) return;
    int const a = en0[e], b = en1[e], c = en2[e];
    bool const to_be_updated = !((nflags[a] | nflags[b] | nflags[c]) & NF_NEUMANN);

    double const xa = x[a] + 1. * UM[a], ya = y[a] + 1. * UM[a + nn];
    double const xb = x[b] + 1. * UM[b], yb = y[b] + 1. * UM[b + nn];
    double const xc = x[c] + 1. * UM[c], yc = y[c] + 1. * UM[c + nn];
    double jac = (xb - xa) * (yc - ya);
    jac -= (xc - xa) * (yb - ya);

    double del_ci = 0.;
    double const surface_old = surface[e];
    double const surface_new = 0.5 * fabs(jac);
    surface[e] = surface_new;
    double cc = conc[e], hh = thick[e], hs = snow[e], hmyi = thick_myi[e], cmyi = conc_myi[e], rr = ridge_ratio[e];
    double cy = 0., hy = 0., hsy = 0.;
    bool const young = K.young_ice != 0;
    if (young) { cy = conc_y[e]; hy = h_y[e]; hsy = hs_y[e]; }
    double const old_conc = cc;
    if ((cc > 0.) && to_be_updated) {
        double const surf_ratio = surface_old / surface_new;
        cc *= surf_ratio;  hh *= surf_ratio;  hs *= surf_ratio;  hmyi *= surf_ratio;
        sig0[e] *= surf_ratio;  sig1[e] *= surf_ratio;  sig2[e] *= surf_ratio;
        rr = 1. - (1. - rr) * fmin(1., cc) / (old_conc * surf_ratio);
        if (young) { hy *= surf_ratio;  cy *= surf_ratio;  hsy *= surf_ratio; }
        if (K.equal_ridging) {
            double const conc_ratio = fmin(1., cc) / old_conc;
            cmyi *= conc_ratio;
            del_ci = 0.;
        } else {
            cmyi *= surf_ratio;
            del_ci = -cmyi;
            cmyi = fmin(cmyi, 1.);
            del_ci += cmyi;
        }
        del_ci *= DAYS_IN_SEC / K.dtime_step;
    }
    double open_water = 1. - cc;
    if (young) open_water -= cy;
    open_water = (open_water < 0.) ? 0. : open_water;
    open_water = (open_water > 1.) ? 1. : open_water;

    double new_conc_young = 0., del_c = 0.;
    if (young) {
        if (cy > 0.) {
            new_conc_young = fmin(1., fmax(0., 1. - cc - open_water));
            if ((cc > K.min_c) && (hh > K.min_h) && (new_conc_young < cy)) {
                double const new_h_young = new_conc_young * hy / cy;
                double const new_hs_young = new_conc_young * hsy / cy;
                double const newice = hy - new_h_young;
                del_c = (cy - new_conc_young) / 10.;
                double const newsnow = hsy - new_hs_young;
                hy = new_h_young;
                hsy = new_hs_young;
                rr = 1. - (1. - rr) * hh / (hh + newice);
                hh += newice;
                hs += newsnow;
            }
        } else {
            hy = 0.;
            hsy = 0.;
        }
    }
    cc = fmin(1., fmax(0., 1. - new_conc_young - open_water + del_c));
    if (young) {
        new_conc_young = fmax(0., fmin(new_conc_young, 1. - cc));
        cy = new_conc_young;
    }
    if (cc > 0.) {
        double test_h_thick = hh / cc;
        test_h_thick = (test_h_thick > 50.) ? 50. : test_h_thick;
        cc = fmin(1. - new_conc_young, hh / test_h_thick);
    } else {
        rr = 0.;  hh = 0.;  hs = 0.;
    }
    cc = (cc > 0.) ? cc : 0.;
    hh = (hh > 0.) ? hh : 0.;
    hmyi = (hmyi > 0.) ? hmyi : 0.;
    hs = (hs > 0.) ? hs : 0.;
    del_ci = -cmyi;
    if (K.myi_with_young) cmyi = fmax(0., fmin(cmyi, cc + cy));
    else cmyi = fmax(0., fmin(cmyi, cc));
    del_ci += cmyi;

    conc[e] = cc;  thick[e] = hh;  snow[e] = hs;  thick_myi[e] = hmyi;  conc_myi[e] = cmyi;  ridge_ratio[e] = rr;
    if (young) { conc_y[e] = cy;  h_y[e] = hy;  hs_y[e] = hsy; }
    del_ci_ridge_myi[e] = del_ci;
}

// ---------------------------------------------------------------------------------------------------
// checkFieldsFast-style device reduction (FE.cpp:14536-14655)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_check(int nn, int ndof, int ne, const double* __restrict__ VT,
        const double* __restrict__ sig0, const double* __restrict__ sig1, const double* __restrict__ sig2,
        const double* __restrict__ damage, const double* __restrict__ conc, const double* __restrict__ thick,
        int* __restrict__ out_i, unsigned long long* __restrict__ out_maxspeed_bits)
{
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    int n_nan = 0, n_speed = 0, n_range = 0;
    double sp = 0.;
    if (t < nn) {
        double const u = VT[t], v = VT[t + nn];
        if (!isfinite(u) || !isfinite(v)) n_nan++;
        else if (t < ndof) { sp = hypot(u, v); if (sp > 5.) n_speed++; }
    }
    if (t < ne) {
        double const d = damage[t], c = conc[t], h = thick[t];
        if (!isfinite(sig0[t]) || !isfinite(sig1[t]) || !isfinite(sig2[t]) || !isfinite(d) || !isfinite(c) || !isfinite(h)) n_nan++;
        else if (d < 0. || d > 1. || c < 0. || c > 1. || h < 0.) n_range++;
    }
    // warp reduce then one atomic per warp (integers and a non-negative double compared as bits)
    for (int o = 16; o > 0; o >>= 1) {
        n_nan += __shfl_down_sync(0xffffffffu, n_nan, o);
        n_speed += __shfl_down_sync(0xffffffffu, n_speed, o);
        n_range += __shfl_down_sync(0xffffffffu, n_range, o);
        sp = fmax(sp, __shfl_down_sync(0xffffffffu, sp, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_nan) atomicAdd(&out_i[0], n_nan);
        if (n_speed) atomicAdd(&out_i[1], n_speed);
        if (n_range) atomicAdd(&out_i[2], n_range);
        atomicMax(out_maxspeed_bits, (unsigned long long)__double_as_longlong(sp));
    }
}

// host numbering <-> internal numbering (nsx_upload / nsx_download); `planes` consecutive planes of n entries
__global__ void __launch_bounds__(TPB)
k_permute_in(int n, int planes, const int* __restrict__ perm, const double* __restrict__ src, double* __restrict__ dst)
{
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int const q = perm[t];
    for (int p = 0; p < planes; ++p) dst[(size_t)p * n + q] = src[(size_t)p * n + t];
}
__global__ void __launch_bounds__(TPB)
k_permute_out(int n, int planes, const int* __restrict__ perm, const double* __restrict__ src, double* __restrict__ dst)
{
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int const q = perm[t];
    for (int p = 0; p < planes; ++p) dst[(size_t)p * n + t] = src[(size_t)p * n + q];
}
// All fields of one nsx_upload / nsx_download call in ONE launch: blockIdx.y selects the field.  Host order lives in
// the transfer arena (one contiguous device buffer, one PCIe copy per field back to back, no per-field kernel).
struct XferEnt { const double* src; double* dst; int n; int planes; int elem; int pad_; };
constexpr int XFER_MAX = 40;
struct XferTable { int count; int pad_; XferEnt e[XFER_MAX]; };
template <int IN>
__global__ void __launch_bounds__(TPB)
k_permute_all(XferTable T, const int* __restrict__ node_perm, const int* __restrict__ elem_perm)
{
    XferEnt const& f = T.e[blockIdx.y];
    int const n = f.n;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        int const q = f.elem ? elem_perm[t] : node_perm[t];
        for (int p = 0; p < f.planes; ++p) {
            if (IN) f.dst[(size_t)p * n + q] = f.src[(size_t)p * n + t];
            else f.dst[(size_t)p * n + t] = f.src[(size_t)p * n + q];
        }
    }
}
// M_shape_coeff[cpt][k] (element-major, reference numbering) from the internal SoA planes
__global__ void __launch_bounds__(TPB)
k_shape_out(int ne, const int* __restrict__ perm, const double* __restrict__ soa, double* __restrict__ aos)
{
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * ne) return;
    int const e = t / 6, k = t - 6 * e;
    aos[t] = soa[(size_t)k * ne + perm[e]];
}

// ---------------------------------------------------------------------------------------------------
// halo exchange (replaces FE.cpp:13963-13996).  ONE kernel per exchange:
//   1. every owner stores (u,v) of its shared nodes straight into the holders' ghost slots of the holders'
//      VT buffer (local memory for in-process groups, NVLink peer memory mapped through CUDA IPC otherwise);
//   2. the last block to finish publishes the new epoch in every holder's flag slot (release, system scope)
//   3. and then waits (bounded spin) until every owner of MY ghosts has published the same epoch.
// The epoch lives in device memory so a captured CUDA graph can be replayed.
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(TPB)
k_halo_exchange(HaloArgs a, int nn_src, const int* __restrict__ src_idx, const int* __restrict__ dst_idx,
                const double* __restrict__ VTsrc, const unsigned long long* my_flags,
                unsigned long long* epoch_ctr, unsigned int* done_ctr, long long max_spins, int* err)
{
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.n_total) {
        int p = 0;
        while (t >= a.peer_begin[p + 1]) ++p;
        int const s = src_idx[t], d = dst_idx[t];
        double* dst = a.peer_vt[p];
        dst[d] = VTsrc[s];
        dst[d + a.peer_nn[p]] = VTsrc[s + nn_src];
    }
    if (!a.sync) return;
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(done_ctr, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    unsigned long long const epoch = *((volatile unsigned long long*)epoch_ctr) + 1ULL;
    flag_signal_wait(a, (int)threadIdx.x, true, epoch, epoch, my_flags, max_spins, err);
    __syncthreads();
    if (threadIdx.x == 0) { *epoch_ctr = epoch; *done_ctr = 0u; }
}

// Mailbox exchange of the tile / direct paths, receiving side: one thread per ghost node polls its mailbox slot for the
// exchange `X.ex` (the owner stored it from its own sub-cycle / sweep kernel), writes it into VT and moves the ghost with
// it (FE.cpp:10539-10553; dt = 0 for smoother sweeps).  Launched AFTER the kernel that issued this rank's pushes, so two
// ranks can never wait for each other's unissued pushes.
__global__ void __launch_bounds__(TPB)
k_ghost_import(MbExchange X, int nn, int ndof, double dt, double* __restrict__ VT, double* __restrict__ disp)
{
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nn - ndof) return;
    double u, v;
    mbx_import(X, t, u, v);
    int const n = ndof + t;
    VT[n] = u;  VT[n + nn] = v;
    if (dt != 0.) { disp[n] += dt * u;  disp[n + nn] += dt * v; }
}
// One smoother sweep with the mailbox exchange in ONE launch (one process per GPU, tile / direct paths).  Blocks
// [0, sweep_blocks) relax the open-water list (FE.cpp:10580-10608): ghost neighbours are read from the mailbox of the
// previous sweep (polling its tag), open-water nodes with holders push their new value.  Blocks behind them re-send the
// sent nodes that are NOT open water (unchanged values) so that every ghost slot carries the tag its holder polls for.
__global__ void __launch_bounds__(TPB)
k_ow_sweep_mb(MbExchange X, int first, int sweep_blocks, int nn, int ndof, const int* __restrict__ ow_list, const int* __restrict__ ow_count,
              const int* __restrict__ n2n, const int* __restrict__ n2n_deg, const uint8_t* __restrict__ nflags,
              const double* __restrict__ node_mass, const double* VTin, double* VTout,
              int n_send, const int* __restrict__ send_src, const int2* __restrict__ send_slot)
{
    if ((int)blockIdx.x < sweep_blocks) {
        int const cnt = *ow_count;
        MbExchange R = X;                   // reads refer to the previous exchange; the first sweep reads the velocity buffer
        if (first) R.on = 0;
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += sweep_blocks * blockDim.x) {
            int const n = ow_list[t];
            int const deg = n2n_deg[n];
            double su = 0., sv = 0.;
            for (int j = 0; j < deg; ++j) {
                int const q = n2n[(size_t)j * nn + n];
                double u, v;
                mbx_velocity(R, q, ndof, nn, VTin, u, v);
                su += u;
                sv += v;
            }
            su = su / deg;  sv = sv / deg;
            VTout[n] = su;
            VTout[n + nn] = sv;
            if (nflags[n] & NF_BTILE) mbx_push(X, n, su, sv);
        }
        return;
    }
    int const t = ((int)blockIdx.x - sweep_blocks) * blockDim.x + threadIdx.x;
    if (t >= n_send) return;
    int const n = send_src[t];
    if (node_mass[n] == 0. && !(nflags[n] & NF_DIRICHLET)) return;       // open-water node: pushed by its sweep thread
    int2 const pe = send_slot[t];
    mb_store(X.peer_mb[pe.x] + (size_t)(X.ex % 3) * X.peer_nmb[pe.x] + pe.y, VTin[n], VTin[n + nn],
             *X.epoch_ctr + (unsigned long long)X.ex);
}

// sending side of a smoother sweep: every sent node is pushed (a sweep only rewrites open-water nodes, the others are
// re-sent unchanged so that the holder finds the tag it polls for)
__global__ void __launch_bounds__(TPB)
k_push_mb(MbExchange X, int nn, int ndof, const double* __restrict__ VT)
{
    int const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ndof) mbx_push(X, t, VT[t], VT[t + nn]);            // nodes without holders return at once
}

// Multi-rank smoother step: one Jacobi sweep over the open-water list AND the ghost exchange of the result in a
// single launch.  All CTAs sweep; the last CTA to finish pushes this rank's sent nodes to their holders, publishes
// the epoch and waits for its neighbours (same protocol as k_halo_exchange).
//
// A sweep only rewrites open-water nodes, so a pair of ranks whose send lists contain none has nothing to exchange
// for the whole smoother.  mode 0 (first sweep): full exchange; every rank marks, per neighbour, whether its send
// list holds open-water nodes, publishes that bit with the epoch and records the neighbour's bit.  mode 1 (middle
// sweeps): only pairs with a bit set on either side push, signal and wait.  mode 2 (last sweep): full exchange again
// (values of skipped pairs are unchanged, so re-storing them is idempotent), which re-establishes the one-exchange
// drift bound of the protocol before the next phase.
__global__ void __launch_bounds__(TPB)
k_ow_sweep_exchange(HaloArgs a, int mode, int nn, const int* __restrict__ ow_list, const int* __restrict__ ow_count,
                    const int* __restrict__ n2n, const int* __restrict__ n2n_deg,
                    const double* VTin, double* VTout,
                    const int* __restrict__ src_idx, const int* __restrict__ dst_idx,
                    const int* __restrict__ push_ptr, const int2* __restrict__ push_ent,
                    int* send_ow, int* pair_active, const unsigned long long* my_flags,
                    unsigned long long* epoch_ctr, unsigned int* done_ctr, long long max_spins, int* err)
{
    int const cnt = *ow_count;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += gridDim.x * blockDim.x) {
        int const n = ow_list[t];
        int const deg = n2n_deg[n];
        double su = 0., sv = 0.;
        for (int j = 0; j < deg; ++j) {
            int const q = n2n[(size_t)j * nn + n];
            su += VTin[q];
            sv += VTin[q + nn];
        }
        VTout[n] = su / deg;
        VTout[n + nn] = sv / deg;
        if (mode == 0)                                       // open-water node in a send list: that pair stays active
            for (int q = push_ptr[n]; q < push_ptr[n + 1]; ++q) atomicOr(send_ow + push_ent[q].x, 1);
    }
    __shared__ bool last;
    __shared__ int s_act[32];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(done_ctr, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    int const i = (int)threadIdx.x;
    // per link: do I exchange with this neighbour in this sweep, and my own open-water bit for it
    if (i < 32) s_act[i] = (mode != 1) ? 1 : (i < a.n_link ? __ldcg(pair_active + i) : 0);
    __syncthreads();
    for (int t = threadIdx.x; t < a.n_total; t += blockDim.x) {
        int p = 0;
        while (t >= a.peer_begin[p + 1]) ++p;
        if (!s_act[a.peer_link[p]]) continue;
        int const s = src_idx[t], d = dst_idx[t];
        double* dst = a.peer_vt[p];
        dst[d] = __ldcg(VTout + s);
        dst[d + a.peer_nn[p]] = __ldcg(VTout + s + nn);
    }
    if (!a.sync) { __syncthreads(); if (threadIdx.x == 0) *done_ctr = 0u; return; }
    __threadfence_system();
    __syncthreads();
    unsigned long long const epoch = *((volatile unsigned long long*)epoch_ctr) + 1ULL;
    int my_bit = 0;
    if (i < a.n_link && mode != 2)
        for (int p = 0; p < a.n_peers; ++p)
            if (a.peer_link[p] == i) my_bit = __ldcg(send_ow + p);
    bool const sel = (i < a.n_link) && s_act[i];
    unsigned long long const v = flag_signal_wait(a, i, sel, epoch | (my_bit ? FLAG_OW : 0ULL), epoch, my_flags, max_spins, err);
    if (mode == 0 && i < a.n_link) pair_active[i] = (my_bit || (v & FLAG_OW)) ? 1 : 0;
    __syncthreads();
    if (threadIdx.x == 0) { *epoch_ctr = epoch; *done_ctr = 0u; }
    if (mode == 2 && i < 32) send_ow[i] = 0;                 // ready for the next model step
}

// ---------------------------------------------------------------------------------------------------
// checkRegridding (FE.cpp:8298-8309) and updateIceDiagnostics (FE.cpp:7860-7900) on the
// resident state.  Products and sums are written with explicit round-to-nearest intrinsics (no FMA contraction)
// so the jacobians, the flip test, the divergence and the forcing values are bit-identical to the host code.
// ---------------------------------------------------------------------------------------------------
// order-preserving map double -> u64: integer atomicMin / atomicMax on the keys order like the doubles
__device__ __forceinline__ unsigned long long ordered_key(double v)
{
    unsigned long long const b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

struct MovedTri { double xa, ya, xb, yb, xc, yc; };
// GmshMesh::vertices(indices, M_UM, 1.)  (gmshmesh.cpp:1929-1939)
__device__ __forceinline__ MovedTri moved_triangle(int a, int b, int c, int nn, const double* __restrict__ x,
                                                   const double* __restrict__ y, const double* __restrict__ UM)
{
    MovedTri t;
    t.xa = x[a] + 1. * UM[a];  t.ya = y[a] + 1. * UM[a + nn];
    t.xb = x[b] + 1. * UM[b];  t.yb = y[b] + 1. * UM[b + nn];
    t.xc = x[c] + 1. * UM[c];  t.yc = y[c] + 1. * UM[c + nn];
    return t;
}
// FiniteElement::jacobian (FE.cpp:1613-1618)
__device__ __forceinline__ double tri_jacobian(MovedTri const& t)
{
    return __dsub_rn(__dmul_rn(t.xb - t.xa, t.yc - t.ya), __dmul_rn(t.xc - t.xa, t.yb - t.ya));
}

// keys[0] = min over elements of minAngles() (FE.cpp:1758-1768), keys[1] / keys[2] = min / max jacobian
__global__ void __launch_bounds__(TPB)
k_regrid_check(int ne, int nn, const int* __restrict__ en0, const int* __restrict__ en1, const int* __restrict__ en2,
               const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ UM,
               unsigned long long* __restrict__ keys)
{
    unsigned long long kang = ~0ULL, kmin = ~0ULL, kmax = 0ULL;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
        MovedTri const t = moved_triangle(en0[e], en1[e], en2[e], nn, x, y, UM);
        double s0 = hypot(t.xb - t.xa, t.yb - t.ya);
        double s1 = hypot(t.xc - t.xb, t.yc - t.yb);
        double s2 = hypot(t.xc - t.xa, t.yc - t.ya);
        double w;                                         // std::sort of three
        if (s0 > s1) { w = s0; s0 = s1; s1 = w; }
        if (s1 > s2) { w = s1; s1 = s2; s2 = w; }
        if (s0 > s1) { w = s0; s0 = s1; s1 = w; }
        double const num = __dsub_rn(__dadd_rn(__dmul_rn(s1, s1), __dmul_rn(s2, s2)), __dmul_rn(s0, s0));
        double const den = __dmul_rn(__dmul_rn(2., s1), s2);
        double ang = acos(__ddiv_rn(num, den));
        ang = __ddiv_rn(__dmul_rn(ang, 45.0), 0.78539816339744828);       // *45/atan(1)
        if (ang == ang) kang = min(kang, ordered_key(ang));               // std::min_element never selects a NaN
        double const jac = tri_jacobian(t);
        kmin = min(kmin, ordered_key(jac));
        kmax = max(kmax, ordered_key(jac));
    }
    for (int o = 16; o > 0; o >>= 1) {
        kang = min(kang, __shfl_down_sync(0xffffffffu, kang, o));
        kmin = min(kmin, __shfl_down_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_down_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(keys + 0, kang);
        atomicMin(keys + 1, kmin);
        atomicMax(keys + 2, kmax);
    }
}

// diag planes: D_conc, D_thick, D_snow_thick, D_sigma[0], D_sigma[1], D_divergence
__global__ void __launch_bounds__(TPB)
k_ice_diagnostics(int ne, int nn, int young_ice,
                  const int* __restrict__ en0, const int* __restrict__ en1, const int* __restrict__ en2,
                  const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ UM,
                  const double* __restrict__ VT,
                  const double* __restrict__ conc, const double* __restrict__ thick, const double* __restrict__ snow,
                  const double* __restrict__ conc_y, const double* __restrict__ h_y, const double* __restrict__ hs_y,
                  const double* __restrict__ sig0, const double* __restrict__ sig1, const double* __restrict__ sig2,
                  double* __restrict__ diag)
{
    int const e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    double dc = conc[e], dh = thick[e], dhs = snow[e];
    if (young_ice) { dc += conc_y[e];  dh += h_y[e];  dhs += hs_y[e]; }
    double const a0 = sig0[e], a1 = sig1[e], a2 = sig2[e];
    int const n[3] = {en0[e], en1[e], en2[e]};
    MovedTri const t = moved_triangle(n[0], n[1], n[2], nn, x, y, UM);
    double const jac = tri_jacobian(t);
    double const vx[3] = {t.xa, t.xb, t.xc}, vy[3] = {t.ya, t.yb, t.yc};
    double div = 0.;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int const kp1 = (k + 1) % 3, kp2 = (k + 2) % 3;
        double const dxN = __ddiv_rn(vy[kp1] - vy[kp2], jac);              // shapeCoeff (FE.cpp:1951-1964)
        double const dyN = __ddiv_rn(vx[kp2] - vx[kp1], jac);
        double const u = VT[n[k]], v = VT[n[k] + nn];
        div = __dadd_rn(div, __dadd_rn(__dmul_rn(dxN, u), __dmul_rn(dyN, v)));
    }
    size_t const E = (size_t)ne;
    diag[e] = dc;
    diag[E + e] = dh;
    diag[2 * E + e] = dhs;
    diag[3 * E + e] = (a0 + a1) / 2.;
    diag[4 * E + e] = hypot((a0 - a1) / 2., a2);
    diag[5 * E + e] = div;
}

// ExternalData::get() for every node (externaldata.cpp:366-436)
__global__ void __launch_bounds__(TPB)
k_forcing_apply(long n, int linear, double c0, double c1, double factor, double bias,
                const double* __restrict__ d0, const double* __restrict__ d1, double* __restrict__ out)
{
    long const i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double value;
    if (linear) value = __dmul_rn(factor, __dadd_rn(__dmul_rn(c0, d0[i]), __dmul_rn(c1, d1[i])));
    else value = __dmul_rn(factor, d0[i]);
    out[i] = __dadd_rn(value, bias);
}

} // namespace nsx

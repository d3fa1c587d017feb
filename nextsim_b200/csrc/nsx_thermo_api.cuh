// nsx_thermo_api.cuh -- FiniteElement::thermo(dt) on the device-resident state: the C ABI (included by nsx_api.cu; the
// kernel is in nsx_thermo.cu).
//
// One thread per element runs nsx::thermo::thermo_element() (nsx_thermo.cuh): OWBulkFluxes + IABulkFluxes (old and young
// ice) + the slab loop fused, so every field is read once and written once.  The kernel is HBM-bound: with the default
// options an element reads 34 doubles (6 forcing, 11 ice state, 17 slab / tracer state), its 3 node ids and the wind at
// its nodes (a node is shared by ~6 elements: 8 B per element from DRAM), and writes 55 doubles (11 ice state, 14 slab /
// tracer state, 30 diagnostics) = 724 algorithmic bytes.  Fields live in the handle's internal (Hilbert) element order,
// one plane per reference member, all thermo-only planes in one allocation.
#pragma once
#include "nsx_thermo.cuh"

namespace nsx {

cudaError_t launch_thermo(thermo::Params const& P, thermo::Arrays const& A, cudaStream_t stream);      // nsx_thermo.cu

// the thermo-only planes (everything but the ice state the dynamics owns), in the order of the X-lists
inline int thermo_private_planes()
{
    int n = 1;      // D_pond_fraction
#define X(f) ++n;
    NSX_THERMO_FORCING(X) NSX_THERMO_STATE(X) NSX_THERMO_DIAG(X)
#undef X
    return n;
}

}  // namespace nsx

// device pointers of every thermo field; allocates and zeroes the thermo-only planes on first use
static nsx::thermo::Arrays& thermo_arrays(nsx_solver* S)
{
    using namespace nsx;
    thermo::Arrays& A = S->th;
    if (S->th_planes.p) return A;
    size_t const ne = (size_t)S->ne;
    S->th_planes.alloc((size_t)thermo_private_planes() * ne);
    S->th_planes.zero(S->stream);
    A = thermo::Arrays{};
    A.ne = S->ne; A.nn = S->nn;
    A.en0 = S->en0.p; A.en1 = S->en1.p; A.en2 = S->en2.p;
    size_t k = 0;
#define X(f) A.f = S->th_planes.p + (k++) * ne;
    NSX_THERMO_FORCING(X) NSX_THERMO_STATE(X)
    A.pond_fraction = S->th_planes.p + (k++) * ne;
    NSX_THERMO_DIAG(X)
#undef X
    A.conc = S->conc.p; A.thick = S->thick.p; A.snow_thick = S->snow.p;
    A.conc_young = S->conc_young.p; A.h_young = S->h_young.p; A.hs_young = S->hs_young.p;
    A.ridge_ratio = S->ridge_ratio.p; A.conc_myi = S->conc_myi.p; A.thick_myi = S->thick_myi.p;
    A.drag_ui = S->drag_ui.p; A.drag_ui_young = S->drag_ui_young.p; A.time_relaxation_damage = S->t_heal.p;
    return A;
}

static void thermo_table(nsx_solver* S, int n, const char* const* names, double* const* host, std::vector<FieldMap>& t, bool upload)
{
    if (n < 0 || (n && (!names || !host))) throw std::invalid_argument("nsx_thermo transfer: bad argument");
    nsx::thermo::Arrays& A = thermo_arrays(S);
    for (int k = 0; k < n; ++k) {
        if (!names[k] || !host[k]) throw std::invalid_argument("nsx_thermo transfer: NULL entry");
        bool shared = false;
        double** slot = nsx::thermo::field_slot(A, names[k], &shared);
        if (!slot) throw std::invalid_argument(std::string("nsx_thermo transfer: unknown field ") + names[k]);
        if (shared && upload)
            throw std::invalid_argument(std::string("nsx_thermo_upload: ") + names[k] + " is part of the ice state, use nsx_upload (NsxFields)");
        t.push_back({host[k], *slot, ELEM});
    }
}

extern "C" void nsx_thermo_params_defaults(NsxThermoParams* p)
{
    if (p) nsx::thermo::params_defaults(*p);
}

extern "C" int nsx_thermo_upload_many(nsx_handle S, int n, const char* const* names, const double* const* host)
{
    NSX_API_BEGIN(S)
    std::vector<FieldMap> t;
    thermo_table(S, n, names, (double* const*)host, t, true);
    std::vector<size_t> off;
    xfer_layout(S, t, off);
    for (size_t k = 0; k < t.size(); ++k)
        NSX_CUDA(cudaMemcpyAsync(S->arena.p + off[k], t[k].host, (size_t)S->ne * sizeof(double), cudaMemcpyHostToDevice, S->stream));
    if (!t.empty()) xfer_permute<1>(S, t, off);
    NSX_CUDA(cudaStreamSynchronize(S->stream));
    NSX_API_END(S)
}

extern "C" int nsx_thermo_download_many(nsx_handle S, int n, const char* const* names, double* const* host)
{
    NSX_API_BEGIN(S)
    std::vector<FieldMap> t;
    thermo_table(S, n, names, host, t, false);
    std::vector<size_t> off;
    xfer_layout(S, t, off);
    if (!t.empty()) xfer_permute<0>(S, t, off);
    for (size_t k = 0; k < t.size(); ++k)
        NSX_CUDA(cudaMemcpyAsync(t[k].host, S->arena.p + off[k], (size_t)S->ne * sizeof(double), cudaMemcpyDeviceToHost, S->stream));
    NSX_CUDA(cudaStreamSynchronize(S->stream));
    NSX_API_END(S)
}

extern "C" int nsx_thermo_upload(nsx_handle S, const char* name, const double* host)
{
    return nsx_thermo_upload_many(S, 1, &name, &host);
}
extern "C" int nsx_thermo_download(nsx_handle S, const char* name, double* host)
{
    return nsx_thermo_download_many(S, 1, &name, &host);
}

// the element forcing of thermo() is ExternalData too (FE.cpp:8063 checkReloadDatasets): same two-slice time interpolation
// as nsx_forcing_load / nsx_forcing_apply, evaluated into the resident M_tair, M_mslp, ...
static double* thermo_forcing_plane(nsx_solver* S, const char* name, const char* who)
{
    if (!name) throw std::invalid_argument(std::string(who) + ": NULL name");
    nsx::thermo::Arrays& A = thermo_arrays(S);
    double** slot = nullptr;
#define X(f) if (!slot && std::string(name) == "M_" #f) slot = &A.f;
    NSX_THERMO_FORCING(X)
#undef X
    if (!slot || std::string(name) == "M_conc_upd")
        throw std::invalid_argument(std::string(who) + ": " + name + " is not an ExternalData forcing variable of thermo()");
    return *slot;
}

extern "C" int nsx_thermo_forcing_load(nsx_handle S, const char* name, int slot, const double* data)
{
    NSX_API_BEGIN(S)
    thermo_forcing_plane(S, name, "nsx_thermo_forcing_load");
    if (slot < 0 || slot > 1 || !data) throw std::invalid_argument("nsx_thermo_forcing_load: bad argument");
    auto& f = S->th_forcing[name];
    if (!f) f.reset(new nsx_solver::ThermoForcing());
    size_t const ne = (size_t)S->ne;
    if (!f->d[slot].p) f->d[slot].alloc(ne);
    if (S->arena.n < ne) { NSX_CUDA(cudaStreamSynchronize(S->stream)); S->arena.alloc(ne); }
    NSX_CUDA(cudaMemcpyAsync(S->arena.p, data, ne * sizeof(double), cudaMemcpyHostToDevice, S->stream));
    std::vector<FieldMap> t{{const_cast<double*>(data), f->d[slot].p, ELEM}};
    std::vector<size_t> off{0};
    xfer_permute<1>(S, t, off);
    NSX_CUDA(cudaStreamSynchronize(S->stream));             // the caller may reuse `data` right away
    f->loaded[slot] = true;
    NSX_API_END(S)
}

extern "C" int nsx_thermo_forcing_apply(nsx_handle S, const char* name, int interp_linear_time, double current_time, double ftime0,
                                        double ftime1, double factor, double bias_correction)
{
    NSX_API_BEGIN(S)
    double* const dst = thermo_forcing_plane(S, name, "nsx_thermo_forcing_apply");
    auto it = S->th_forcing.find(name);
    if (it == S->th_forcing.end() || !it->second->loaded[0] || (interp_linear_time && !it->second->loaded[1]))
        throw std::runtime_error("nsx_thermo_forcing_apply: time slice not loaded (nsx_thermo_forcing_load)");
    double c0 = 1., c1 = 0.;
    if (interp_linear_time) {                               // externaldata.cpp:368-370
        double const fdt = std::fabs(ftime1 - ftime0);
        c0 = std::fabs(current_time - ftime1) / fdt;
        c1 = std::fabs(current_time - ftime0) / fdt;
    }
    double const* const d0 = it->second->d[0].p;
    double const* const d1 = interp_linear_time ? it->second->d[1].p : d0;
    k_forcing_apply<<<nblk(S->ne), TPB, 0, S->stream>>>((long)S->ne, interp_linear_time, c0, c1, factor, bias_correction, d0, d1, dst);
    NSX_CUDA(cudaGetLastError());
    NSX_API_END(S)
}

extern "C" int nsx_thermo(nsx_handle S, const NsxThermoParams* p, int dt, double current_time)
{
    NSX_API_BEGIN(S)
    if (!p) throw std::invalid_argument("nsx_thermo: NULL parameters");
    if (const char* e = nsx::thermo::validate(*p, dt)) throw std::invalid_argument(e);
    nsx::thermo::Arrays A = thermo_arrays(S);
    A.wind = S->wind.p; A.ocean = S->ocean.p; A.VT = S->VT[S->cur];
    nsx::thermo::Params const P = nsx::thermo::make_params(*p, dt, current_time);
    if (S->ne > 0) {
        NSX_CUDA(nsx::launch_thermo(P, A, S->stream));
        ++S->n_thermo_launch;
    }
    NSX_API_END(S)
}

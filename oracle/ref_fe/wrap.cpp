// wrap.cpp -- C harness around the reference's own function bodies (TEST INFRASTRUCTURE ONLY, oracle/).
//
// Compiles the text cut from /root/reference/model/finiteelement.cpp by extract.py (explicitSolve, update,
// updateSigmaDamage, updateSigmaVP/EVP/MEVP, updateGhosts, geometry helpers, checkRegridding,
// updateIceDiagnostics, ...) against stub_fe.hpp and exposes set / run / get entry points for ctypes.  The harness
// only fills members and calls the reference functions; P ranks run explicitSolve() on P threads so that the
// reference's updateGhosts() exchanges through the in-process Communicator stand-in.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "stub_fe.hpp"
#include "../../include/nsx.h"          // NsxThermoParams: the option block both sides are driven with

namespace Nextsim {
#include "ref_fe_bodies.inc"
}

using Nextsim::FiniteElement;

namespace {
struct Rank {
    FiniteElement fe;
    Nextsim::BamgMeshStub bamg;
    std::vector<double> nec, nc;
};
struct Harness {
    Nextsim::World world;
    std::vector<std::unique_ptr<Rank>> ranks;
    std::string err;
};

// mirror of oracle.OrcParams / NsxDynParams (same field order as oracle/oracle.py)
struct Params {
    int dynamics_type, basal_stress_type, ice_cat_type, substeps, equal_ridging, newice_type,
        use_young_ice_in_myi_reset, stop_after_substeps, skip_ow_smoother, pad_;
    double dtime_step, ocean_turning_angle_rad, min_h, min_c, young, nu0, tan_phi, compr_strength, compaction_param,
        undamaged_time_relaxation_sigma, exponent_relaxation_sigma, compression_factor, exponent_compression_factor,
        quad_drag_coef_water, evp_e, evp_Pstar, evp_C, evp_dmin, mevp_alpha, mevp_beta, basal_k1, basal_k2, basal_Cb,
        basal_u0;
};

std::vector<double>* field(FiniteElement& fe, std::string const& n)
{
    if (n == "M_VT") return &fe.M_VT;
    if (n == "M_UM") return &fe.M_UM;
    if (n == "M_UT") return &fe.M_UT;
    if (n == "M_wind") return &fe.M_wind.data;
    if (n == "M_ocean") return &fe.M_ocean.data;
    if (n == "M_ssh") return &fe.M_ssh.data;
    if (n == "M_element_depth") return &fe.M_element_depth.data;
    if (n == "lat") return &fe.M_mesh.M_lat;
    if (n == "M_surface") return &fe.M_surface;
    if (n == "M_delta_x") return &fe.M_delta_x;
    if (n == "M_sigma0") return &fe.M_sigma[0];
    if (n == "M_sigma1") return &fe.M_sigma[1];
    if (n == "M_sigma2") return &fe.M_sigma[2];
    if (n == "D_sigma0") return &fe.D_sigma[0];
    if (n == "D_sigma1") return &fe.D_sigma[1];
#define F(x) if (n == #x) return &fe.x;
    F(M_conc) F(M_thick) F(M_snow_thick) F(M_damage) F(M_ridge_ratio) F(M_conc_young) F(M_h_young) F(M_hs_young)
    F(M_thick_myi) F(M_conc_myi) F(M_Cohesion) F(M_time_relaxation_damage) F(M_drag_ui) F(M_drag_ui_young)
    F(M_random_number) F(D_tau_a) F(D_tau_w) F(D_del_ci_ridge_myi) F(D_conc) F(D_thick) F(D_snow_thick) F(D_divergence)
    // thermo(): state and diagnostics
    F(M_sst) F(M_sss) F(M_tsurf_young) F(M_conc_upd) F(M_del_vi_tend) F(M_freeze_days) F(M_freeze_onset) F(M_conc_summer)
    F(M_thick_summer) F(M_fyi_fraction) F(M_age_det) F(M_age) F(M_pond_volume) F(M_lid_volume) F(M_drag_ti) F(M_drag_ti_young)
    F(D_tau_ow) F(D_pond_fraction) F(D_Qa) F(D_Qsw) F(D_Qlw) F(D_Qsh) F(D_Qlh) F(D_Qo) F(D_Qnosun) F(D_Qsw_ocean) F(D_Qassim)
    F(D_delS) F(D_fwflux_ice) F(D_fwflux) F(D_brine) F(D_evap) F(D_rain) F(D_vice_melt) F(D_del_vi_young) F(D_del_hi)
    F(D_del_hi_young) F(D_newice) F(D_mlt_top) F(D_mlt_bot) F(D_snow2ice) F(D_albedo) F(D_sialb) F(D_del_ci_mlt_myi)
    F(D_del_vi_mlt_myi) F(D_del_ci_rplnt_myi) F(D_del_vi_rplnt_myi)
#undef F
    // thermo(): forcing (ExternalData)
#define X(x) if (n == #x) return &fe.x.data;
    X(M_tair) X(M_mixrat) X(M_dair) X(M_sphuma) X(M_mslp) X(M_Qsw_in) X(M_Qlw_in) X(M_tcc) X(M_precip) X(M_snowfall) X(M_snowfr)
    X(M_mld) X(M_ocean_temp) X(M_ocean_salt)
#undef X
    if (n == "M_tice0") return &fe.M_tice[0];
    if (n == "M_tice1") return fe.M_tice.size() > 1 ? &fe.M_tice[1] : nullptr;
    if (n == "M_tice2") return fe.M_tice.size() > 2 ? &fe.M_tice[2] : nullptr;
    return nullptr;
}
}  // namespace

extern "C" {

void* ref_fe_create(int nranks)
{
    Harness* H = new Harness();
    H->world.n = nranks;
    for (int r = 0; r < nranks; ++r) {
        H->ranks.emplace_back(new Rank());
        FiniteElement& fe = H->ranks.back()->fe;
        fe.M_comm.world = &H->world;
        fe.M_comm.me = r;
        fe.M_rank = r;
        fe.M_sigma.resize(3);
        fe.D_sigma.resize(2);
        fe.M_tice.resize(1);
        fe.bamgmesh = &H->ranks.back()->bamg;
        fe.M_extract_local_index.resize(nranks);
        fe.M_local_ghosts_local_index.resize(nranks);
    }
    return H;
}
void ref_fe_destroy(void* h) { delete (Harness*)h; }
const char* ref_fe_last_error(void* h) { return ((Harness*)h)->err.c_str(); }

// local mesh of one rank: coordinates, 1-based triangles (owned first), masks
int ref_fe_set_mesh(void* h, int r, int nn, int ndof, const double* x, const double* y, int ne, const int* tri1,
                    const unsigned char* mask_dirichlet, int n_neumann, const int* neumann_flags)
{
    FiniteElement& fe = ((Harness*)h)->ranks[r]->fe;
    fe.M_num_nodes = nn; fe.M_local_ndof = ndof; fe.M_num_elements = ne;
    fe.M_mesh.M_num_nodes = nn;
    fe.M_mesh.M_nodes.clear();
    for (int i = 0; i < nn; ++i) {
        Nextsim::entities::GMSHPoint p;
        p.id = i + 1;
        p.coords = {x[i], y[i]};
        fe.M_mesh.M_nodes[i + 1] = p;
    }
    fe.M_elements.assign(ne, FiniteElement::element_type());
    for (int e = 0; e < ne; ++e) {
        auto& el = fe.M_elements[e];
        el.number = e + 1;
        el.indices = {tri1[3 * e], tri1[3 * e + 1], tri1[3 * e + 2]};
        el.ghostNodes.assign(3, false);
        for (int i = 0; i < 3; ++i) el.ghostNodes[i] = (el.indices[i] - 1 >= ndof);     // gmshmesh.cpp:1298-1301
    }
    fe.M_mesh.M_triangles = fe.M_elements;
    fe.M_mask_dirichlet.assign(nn, false);
    for (int i = 0; i < nn; ++i) fe.M_mask_dirichlet[i] = mask_dirichlet[i] != 0;
    // M_neumann_flags: sorted node ids; M_neumann_nodes: both dofs (FE.cpp:236-262)
    fe.M_neumann_flags.assign(neumann_flags, neumann_flags + n_neumann);
    fe.M_neumann_nodes.clear();
    for (int k = 0; k < n_neumann; ++k) {
        fe.M_neumann_nodes.push_back(neumann_flags[k]);
        fe.M_neumann_nodes.push_back(neumann_flags[k] + nn);
    }
    for (const char* n : {"M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "D_tau_a", "D_tau_w"}) field(fe, n)->assign(2 * (size_t)nn, 0.);
    fe.M_ssh.data.assign(nn, 0.);
    for (const char* n : {"M_conc", "M_thick", "M_snow_thick", "M_damage", "M_ridge_ratio", "M_conc_young", "M_h_young",
                          "M_hs_young", "M_thick_myi", "M_conc_myi", "M_Cohesion", "M_time_relaxation_damage", "M_drag_ui",
                          "M_drag_ui_young", "M_random_number", "D_del_ci_ridge_myi", "M_sigma0", "M_sigma1", "M_sigma2",
                          "M_element_depth", "M_surface", "M_delta_x", "D_divergence"})
        field(fe, n)->assign(ne, 0.);
    fe.M_sst.assign(ne, 0.); fe.M_sss.assign(ne, 0.); fe.M_tsurf_young.assign(ne, 0.); fe.M_tice[0].assign(ne, 0.);
    fe.D_dmean.assign(ne, 0.); fe.D_dmax.assign(ne, 0.);
    return 0;
}

// bamg tables as BamgConvertMeshx leaves them (row-major doubles, NaN padding / count in the last column)
int ref_fe_set_bamg(void* h, int r, const double* nec, int nec_w, const double* nc, int nc_w)
{
    Rank& R = *((Harness*)h)->ranks[r];
    int const nn = R.fe.M_num_nodes;
    R.nec.assign(nec, nec + (size_t)nn * nec_w);
    R.nc.assign(nc, nc + (size_t)nn * nc_w);
    R.bamg.NodalElementConnectivitySize[0] = nn; R.bamg.NodalElementConnectivitySize[1] = nec_w;
    R.bamg.NodalElementConnectivity = R.nec.data();
    R.bamg.NodalConnectivitySize[0] = nn; R.bamg.NodalConnectivitySize[1] = nc_w;
    R.bamg.NodalConnectivity = R.nc.data();
    return 0;
}

// which = 0: M_extract_local_index[proc] (what I send to proc), 1: M_local_ghosts_local_index[proc] (what proc fills)
int ref_fe_set_halo(void* h, int r, int which, int proc, const int* idx, int n)
{
    FiniteElement& fe = ((Harness*)h)->ranks[r]->fe;
    auto& lists = which == 0 ? fe.M_extract_local_index : fe.M_local_ghosts_local_index;
    auto& procs = which == 0 ? fe.M_recipients_proc_id : fe.M_local_ghosts_proc_id;
    lists[proc].assign(idx, idx + n);
    if (n && std::find(procs.begin(), procs.end(), proc) == procs.end()) procs.push_back(proc);
    std::sort(procs.begin(), procs.end());
    return 0;
}

int ref_fe_set_double(void* h, int r, const char* name, const double* v, long n)
{
    Harness* H = (Harness*)h;
    auto* f = field(H->ranks[r]->fe, name);
    if (!f) { H->err = std::string("unknown field ") + name; return 2; }
    f->assign(v, v + n);
    return 0;
}
long ref_fe_size_double(void* h, int r, const char* name)
{
    auto* f = field(((Harness*)h)->ranks[r]->fe, name);
    return f ? (long)f->size() : -1;
}
int ref_fe_get_double(void* h, int r, const char* name, double* out, long n)
{
    Harness* H = (Harness*)h;
    auto* f = field(H->ranks[r]->fe, name);
    if (!f || (long)f->size() != n) { H->err = std::string("bad get ") + name; return 2; }
    std::memcpy(out, f->data(), sizeof(double) * n);
    return 0;
}
// M_shape_coeff[cpt][k], element-major
int ref_fe_get_shape_coeff(void* h, int r, double* out)
{
    FiniteElement& fe = ((Harness*)h)->ranks[r]->fe;
    for (size_t e = 0; e < fe.M_shape_coeff.size(); ++e)
        for (int k = 0; k < 6; ++k) out[6 * e + k] = fe.M_shape_coeff[e][k];
    return 0;
}

int ref_fe_set_params(void* h, const Params* p, double regrid_angle)
{
    Harness* H = (Harness*)h;
    for (auto& R : H->ranks) {
        FiniteElement& fe = R->fe;
        // the option names of model/options.cpp the cut bodies look up
        fe.vm.set("dynamics.substeps", p->substeps);
        fe.vm.set("dynamics.min_h", p->min_h);
        fe.vm.set("dynamics.min_c", p->min_c);
        fe.vm.set("dynamics.undamaged_time_relaxation_sigma", p->undamaged_time_relaxation_sigma);
        fe.vm.set("dynamics.exponent_relaxation_sigma", p->exponent_relaxation_sigma);
        fe.vm.set("dynamics.evp.e", p->evp_e);
        fe.vm.set("dynamics.evp.Pstar", p->evp_Pstar);
        fe.vm.set("dynamics.evp.C", p->evp_C);
        fe.vm.set("dynamics.evp.dmin", p->evp_dmin);
        fe.vm.set("dynamics.mevp.alpha", p->mevp_alpha);
        fe.vm.set("dynamics.mevp.beta", p->mevp_beta);
        fe.vm.set("dynamics.Lemieux_basal_k1", p->basal_k1);
        fe.vm.set("dynamics.Lemieux_basal_k2", p->basal_k2);
        fe.vm.set("dynamics.Lemieux_basal_Cb", p->basal_Cb);
        fe.vm.set("dynamics.Lemieux_basal_u_0", p->basal_u0);
        fe.vm.set("thermo.diffusivity_sst", 0.);
        fe.vm.set("thermo.diffusivity_sss", 0.);
        fe.vm.set("age.equal_ridging", p->equal_ridging);
        fe.vm.set("thermo.newice_type", p->newice_type);
        fe.vm.set("age.include_young_ice", p->use_young_ice_in_myi_reset);
        fe.vm.set("numerics.regrid_angle", regrid_angle);
        // members set by initOptAndParam() (FE.cpp:1089-1240) from the same options
        fe.dtime_step = p->dtime_step;
        fe.ocean_turning_angle_rad = p->ocean_turning_angle_rad;
        fe.nu0 = p->nu0; fe.young = p->young; fe.compaction_param = p->compaction_param;
        fe.undamaged_time_relaxation_sigma = p->undamaged_time_relaxation_sigma;
        fe.exponent_relaxation_sigma = p->exponent_relaxation_sigma;
        fe.compression_factor = p->compression_factor;
        fe.exponent_compression_factor = p->exponent_compression_factor;
        fe.compr_strength = p->compr_strength; fe.tan_phi = p->tan_phi;
        fe.quad_drag_coef_water = p->quad_drag_coef_water;
        // enum values of NsxDynParams / OrcParams (include/nsx.h:34-37): dynamics 0 bbm, 3 evp, 4 mevp; basal 0 none, 1 lemieux; icecat 0 classic, 1 young
        fe.M_dynamics_type = p->dynamics_type == 0 ? Nextsim::setup::DynamicsType::BBM
                           : p->dynamics_type == 3 ? Nextsim::setup::DynamicsType::EVP : Nextsim::setup::DynamicsType::mEVP;
        fe.M_basal_stress_type = p->basal_stress_type == 1 ? Nextsim::setup::BasalStressType::LEMIEUX : Nextsim::setup::BasalStressType::NONE;
        fe.M_ice_cat_type = p->ice_cat_type == 1 ? Nextsim::setup::IceCategoryType::YOUNG_ICE : Nextsim::setup::IceCategoryType::CLASSIC;
        fe.initFETensors();             // reference text: M_Dunit from nu0
    }
    return 0;
}

int ref_fe_calc_cohesion(void* h, int r, double C_fix, double C_alea)
{
    FiniteElement& fe = ((Harness*)h)->ranks[r]->fe;
    fe.C_fix = C_fix; fe.C_alea = C_alea;
    fe.calcCohesion();
    return 0;
}

// explicitSolve() on every rank, one thread per rank (the reference's updateGhosts blocks on its neighbours)
int ref_fe_explicit_solve(void* h)
{
    Harness* H = (Harness*)h;
    std::vector<std::thread> th;
    std::vector<std::string> errs(H->ranks.size());
    for (size_t r = 0; r < H->ranks.size(); ++r)
        th.emplace_back([H, r, &errs] {
            try { H->ranks[r]->fe.explicitSolve(); } catch (std::exception const& e) { errs[r] = e.what(); }
        });
    for (auto& t : th) t.join();
    for (auto& e : errs) if (!e.empty()) { H->err = e; return 2; }
    return 0;
}

int ref_fe_update(void* h)
{
    Harness* H = (Harness*)h;
    try {
        for (auto& R : H->ranks) { std::vector<double> UM_P = R->fe.M_UM; R->fe.update(UM_P); }     // FE.cpp:8204-8211
    } catch (std::exception const& e) { H->err = e.what(); return 2; }
    return 0;
}

// out: min angle, min jacobian, max jacobian (local part); flags: flip, regrid_local
int ref_fe_check_regridding(void* h, int r, double* out, int* flags)
{
    Harness* H = (Harness*)h;
    FiniteElement& fe = H->ranks[r]->fe;
    try {
        out[0] = fe.minAngle(fe.M_mesh, fe.M_UM, 1., true);
        double mn = 1e300, mx = -1e300;
        for (auto const& el : fe.M_mesh.triangles()) {
            double const j = fe.jacobian(el, fe.M_mesh, fe.M_UM, 1.);
            mn = std::min(mn, j); mx = std::max(mx, j);
        }
        out[1] = mn; out[2] = mx;
        flags[0] = fe.flip(fe.M_mesh, fe.M_UM, 1.) ? 1 : 0;
        flags[1] = fe.checkRegridding() ? 1 : 0;
    } catch (std::exception const& e) { H->err = e.what(); return 2; }
    return 0;
}

int ref_fe_update_ice_diagnostics(void* h)
{
    Harness* H = (Harness*)h;
    try { for (auto& R : H->ranks) R->fe.updateIceDiagnostics(); }
    catch (std::exception const& e) { H->err = e.what(); return 2; }
    return 0;
}

// ---- thermo(): options of NsxThermoParams -> the option map and members initOptAndParam() fills (FE.cpp:1089-1300) ----
int ref_fe_thermo_setup(void* h, const NsxThermoParams* p, double current_time)
{
    Harness* H = (Harness*)h;
    char md[8];
    std::snprintf(md, sizeof md, "%02d%02d", p->reset_month, p->reset_day);
    for (auto& R : H->ranks) {
        FiniteElement& fe = R->fe;
        fe.vm.set("thermo.ocean_nudge_timeT_days", p->ocean_nudge_timeT_days);
        fe.vm.set("thermo.ocean_nudge_timeS_days", p->ocean_nudge_timeS_days);
        fe.vm.set("ideal_simul.constant_Qdw", p->Qdw_const);
        fe.vm.set("ideal_simul.constant_Fdw", p->Fdw_const);
        fe.vm.set("ideal_simul.constant_mld", p->constant_mld);
        fe.vm.set("thermo.hnull", p->hnull);
        fe.vm.set("thermo.PhiF", p->PhiF);
        fe.vm.set("thermo.PhiM", p->PhiM);
        fe.vm.set("thermo.newice_type", p->newice_type);
        fe.vm.set("thermo.melt_type", p->melt_type);
        fe.vm.set("thermo.use_assim_flux", p->use_assim_flux);
        fe.vm.set("thermo.assim_flux_exponent", p->assim_flux_exponent);
        fe.vm.set("dynamics.use_temperature_dependent_healing", p->temp_dep_healing);
        fe.vm.set("thermo.I_0", p->I_0);
        fe.vm.set("age.include_young_ice", p->use_young_ice_in_myi_reset);
        fe.vm.set("age.reset_date", std::string(md));
        fe.vm.set("age.reset_by_date", p->reset_by_date);
        fe.vm.set("age.reset_freeze_days", p->freeze_days_threshold);
        fe.vm.set("age.equal_melting", p->equal_melting);
        fe.vm.set("thermo.use_meltponds", p->use_meltponds);
        fe.vm.set("thermo.meltpond_runoff_fraction", p->meltpond_runoff_fraction);
        fe.vm.set("thermo.meltpond_depth_to_fraction", p->meltpond_depth_to_fraction);
        fe.vm.set("thermo.drag_ocean_t", p->drag_ocean_t);
        fe.vm.set("thermo.drag_ocean_q", p->drag_ocean_q);
        fe.vm.set("thermo.alb_scheme", p->alb_scheme);
        fe.vm.set("thermo.alb_ice", p->alb_ice);
        fe.vm.set("thermo.alb_sn", p->alb_sn);
        fe.vm.set("thermo.alb_ponds", p->alb_ponds);
        fe.vm.set("thermo.force_neutral_atmosphere", p->force_neutral_atmosphere);
        fe.vm.set("thermo.zref_wind", p->zref_wind);
        fe.vm.set("thermo.zref_temp", p->zref_temp);
        fe.vm.set("thermo.limiting_lengthscale", p->limiting_lengthscale);
        fe.dtime_step = p->dtime_step;
        fe.M_current_time = current_time;
        fe.M_ocean_albedo = p->ocean_albedo;
        fe.M_ks = p->ks;
        fe.M_freezingpoint_mu = p->freezingpoint_mu;
        fe.M_Csens_io = p->Csens_io;
        fe.M_flooding = p->flooding != 0;
        fe.time_relaxation_damage = p->time_relaxation_damage;
        fe.deltaT_relaxation_damage = p->deltaT_relaxation_damage;
        fe.h_young_min = p->h_young_min;
        fe.h_young_max_sharp = .5 * (p->h_young_min + p->h_young_max);                  // FE.cpp:1198
        fe.quad_drag_coef_air = p->quad_drag_coef_air;
        fe.M_thermo_type = p->thermo_type == 0 ? Nextsim::setup::ThermoType::ZERO_LAYER : Nextsim::setup::ThermoType::WINTON;
        fe.M_ocean_type = p->ocean_constant ? Nextsim::setup::OceanType::CONSTANT : Nextsim::setup::OceanType::TOPAZ4R;
        fe.M_Qio_type = p->Qio_type == 0 ? Nextsim::setup::OceanHeatfluxScheme::BASIC : Nextsim::setup::OceanHeatfluxScheme::EXCHANGE;
        fe.M_freezingpoint_type = p->freezingpoint_type == 0 ? Nextsim::setup::FreezingPointType::LINEAR : Nextsim::setup::FreezingPointType::UNESCO;
        fe.M_ice_cat_type = p->ice_cat_young ? Nextsim::setup::IceCategoryType::YOUNG_ICE : Nextsim::setup::IceCategoryType::CLASSIC;
        fe.M_sphuma.initialized = p->have_sphuma != 0;
        fe.M_mixrat.initialized = p->have_mixrat != 0;
        fe.M_Qlw_in.initialized = p->have_Qlw_in != 0;
        fe.M_snowfr.initialized = p->have_snowfr != 0;
        fe.M_snowfall.initialized = p->have_snowfall != 0;
        fe.M_mld.initialized = p->have_mld != 0;
        // M_tice has one layer under the zero-layer scheme and three under Winton (FE.cpp:1330-1345)
        size_t const ne = (size_t)fe.M_num_elements;
        fe.M_tice.resize(p->thermo_type == 0 ? 1 : 3);
        for (auto& t : fe.M_tice) t.resize(ne, 0.);
        for (const char* n : {"M_sst", "M_sss", "M_tsurf_young", "M_conc_upd", "M_del_vi_tend", "M_freeze_days", "M_freeze_onset",
                              "M_conc_summer", "M_thick_summer", "M_fyi_fraction", "M_age_det", "M_age", "M_pond_volume",
                              "M_lid_volume", "M_drag_ti", "M_drag_ti_young", "D_tau_ow", "D_pond_fraction", "D_Qa", "D_Qsw",
                              "D_Qlw", "D_Qsh", "D_Qlh", "D_Qo", "D_Qnosun", "D_Qsw_ocean", "D_Qassim", "D_delS", "D_fwflux_ice",
                              "D_fwflux", "D_brine", "D_evap", "D_rain", "D_vice_melt", "D_del_vi_young", "D_del_hi",
                              "D_del_hi_young", "D_newice", "D_mlt_top", "D_mlt_bot", "D_snow2ice", "D_albedo", "D_sialb",
                              "D_del_ci_mlt_myi", "D_del_vi_mlt_myi", "D_del_ci_rplnt_myi", "D_del_vi_rplnt_myi", "M_tair",
                              "M_mixrat", "M_dair", "M_sphuma", "M_mslp", "M_Qsw_in", "M_Qlw_in", "M_tcc", "M_precip",
                              "M_snowfall", "M_snowfr", "M_mld", "M_ocean_temp", "M_ocean_salt"})
            field(fe, n)->resize(ne, 0.);
    }
    return 0;
}

// month * 100 + day of the stand-in for datenumToString(t, "%m%d") the reference bodies are compiled against
int ref_fe_month_day(double datenum) { return std::atoi(Nextsim::datenumToString(datenum, "%m%d").c_str()); }

int ref_fe_thermo(void* h, int dt)
{
    Harness* H = (Harness*)h;
    try { for (auto& R : H->ranks) R->fe.thermo(dt); }
    catch (std::exception const& e) { H->err = e.what(); return 2; }
    return 0;
}

}  // extern "C"

// stub_fe.hpp -- stand-in for the reference's model/finiteelement.hpp, written for this repository.
//
// TEST INFRASTRUCTURE ONLY (oracle/).  The reference class cannot be compiled here (Boost, Gmsh, NetCDF, MPI are
// absent), but the BODIES of its hot-path member functions only need libm and a handful of members.  This header
// declares exactly those members with the reference's names and types, plus minimal stand-ins for the library
// types the bodies mention (program_options map, Timer, LOG, ExternalData, Communicator, boost::mpi::all_reduce),
// so that the function definitions cut verbatim from /root/reference by extract.py compile unchanged.
// Nothing here computes physics: every arithmetic statement that runs is reference text.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "constants.hpp"        // the reference's own model/constants.hpp (plain C++), found through -I
#include "enums.hpp"            // the reference's own model/enums.hpp (plain C++): setup::*, schemes::*

#define PI 3.141592653589793          // contrib/mapx/include/mapx.h:47, reaches finiteelement.cpp through gmshmesh.hpp

// ---- LOG(level) << ... : swallowed ----
namespace Nextsim { struct NullLog { template <class T> NullLog& operator<<(T const&) { return *this; } }; }
#define LOG(level) ::Nextsim::NullLog()

namespace boost { namespace mpi {
template <class T> struct minimum {};
// one rank at a time calls these from the harness: the reduction over ranks is done by the caller
template <class C, class T, class Op> T all_reduce(C const&, T const& v, Op) { return v; }
template <class C> void all_reduce(C const&, bool const& in, bool& out, std::plus<bool>) { out = in; }
}}

namespace Nextsim {

inline double real(int v) { return double(v); }      // FE.cpp:10185 `real(steps)`: std::real(int) -> double

using physical::sigma_sb;        // FE.cpp:6386 names it unqualified

// core/include/date.hpp:87-120 (boost::date_time there): nextsim time = decimal days since 1900-01-01 00:00; the date is the
// epoch + static_cast<long>(t) days, the time of day is rounded to milliseconds and 24:00:00.000 carries into the next date.
// Only the "%m%d" format is used on the path (FE.cpp:5216).  Written independently of the product's month_day(): a table walk
// over the years instead of the era arithmetic, so that the two check each other (tests/test_thermo_cpu.py::test_dates).
inline std::string datenumToString(double const& datenum, std::string const& format)
{
    if (format != "%m%d") throw std::runtime_error("ref_fe stub: datenumToString format " + format);
    long days = static_cast<long>(datenum);
    double const fractionalDay = datenum - std::floor(datenum);
    if (static_cast<long>(std::floor(fractionalDay * 24.0 * 60.0 * 60.0 * 1000.0 + 0.5)) >= 86400000L) ++days;
    if (days < 0) throw std::runtime_error("ref_fe stub: time before 1900-01-01");
    int year = 1900;
    auto leap = [](int y) { return (y % 4 == 0 && y % 100 != 0) || y % 400 == 0; };
    while (days >= (leap(year) ? 366 : 365)) { days -= leap(year) ? 366 : 365; ++year; }
    static const int mlen[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int month = 0;
    while (days >= mlen[month] + (month == 1 && leap(year) ? 1 : 0)) { days -= mlen[month] + (month == 1 && leap(year) ? 1 : 0); ++month; }
    char buf[8];
    std::snprintf(buf, sizeof buf, "%02d%02d", month + 1, (int)days + 1);
    return buf;
}

// ---- boost::program_options::variables_map stand-in ----
struct OptValue {
    double v = 0.;
    std::string s;
    template <class T> T as() const { return static_cast<T>(v); }
};
template <> inline std::string OptValue::as<std::string>() const { return s; }
struct OptMap {
    std::map<std::string, OptValue> m;
    OptValue const& operator[](std::string const& k) const {
        auto it = m.find(k);
        if (it == m.end()) throw std::runtime_error("ref_fe stub: option not set: " + k);
        return it->second;
    }
    int count(std::string const& k) const { return (int)m.count(k); }
    void set(std::string const& k, double v) { m[k].v = v; }
    void set(std::string const& k, std::string const& v) { m[k].s = v; }
};

struct Timer {
    void tick(std::string const&) {}
    void tock(std::string const&) {}
};

// ModelVariable derives from std::vector<double> (model/model_variable.hpp:26)
struct ModelVariable : std::vector<double> { using std::vector<double>::vector; };

// ExternalData: explicitSolve only uses operator[](i) and getVector() (model/externaldata.hpp)
struct ExternalData {
    std::vector<double> data;
    bool initialized = false;               // thermo() branches on isInitialized() (which forcing variables the dataset has)
    bool isInitialized() const { return initialized; }
    double operator[](int i) const { return data[i]; }
    std::vector<double> getVector() const { return data; }
};

namespace entities {
struct GMSHPoint { std::vector<double> coords; int id = 0; };
struct GMSHElement {                        // core/include/entities.hpp:35-152, the members the path reads
    std::vector<int> indices;               // 1-based local node ids
    std::vector<bool> ghostNodes;
    int number = 0;
};
}

// in-process stand-in for Communicator (boost::mpi): blocking mailboxes between the threads that play the ranks
struct World {
    int n = 1;
    std::mutex mu;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<std::vector<double>>> box;       // (src, dst) -> messages
};
struct Communicator {
    World* world = nullptr;
    int me = 0;
    int size() const { return world ? world->n : 1; }
    void barrier() const {}                 // thermo() starts with one; the harness calls it rank by rank
    int rank() const { return me; }
    void send(int dst, int /*tag*/, std::vector<double> const& v) const {
        std::lock_guard<std::mutex> g(world->mu);
        world->box[{me, dst}].push_back(v);
        world->cv.notify_all();
    }
    void recv(int src, int /*tag*/, std::vector<double>& v) const {
        std::unique_lock<std::mutex> g(world->mu);
        auto& q = world->box[{src, me}];
        world->cv.wait(g, [&] { return !q.empty(); });
        v = std::move(q.front());
        q.pop_front();
    }
};

class GmshMesh {
public:
    typedef entities::GMSHPoint point_type;
    typedef entities::GMSHElement element_type;
    std::map<int, point_type> const& nodes() const { return M_nodes; }
    std::vector<element_type> const& triangles() const { return M_triangles; }
    int numTriangles() const { return (int)M_triangles.size(); }
    std::vector<double> lat() const { return M_lat; }       // gmshmesh.cpp:1800-1824 needs mapx: supplied by the harness
    // definitions cut from core/src/gmshmesh.cpp:1918-1939
    std::vector<std::vector<double>> vertices(std::vector<int> const& indices) const;
    std::vector<std::vector<double>> vertices(std::vector<int> const& indices, std::vector<double> const& um, double factor) const;

    mutable std::map<int, point_type> M_nodes;              // `M_nodes[id]` is used inside a const member
    std::vector<element_type> M_triangles;
    std::vector<double> M_lat;
    int M_num_nodes = 0;
};
struct GmshMeshSeq {};                                      // mesh_type_root: only named in overloads that are not cut

struct BamgMeshStub {                                       // contrib/bamg/include/BamgMesh.h: tables are double*
    int NodalElementConnectivitySize[2] = {0, 0};
    double* NodalElementConnectivity = nullptr;
    int NodalConnectivitySize[2] = {0, 0};
    double* NodalConnectivity = nullptr;
};

class FiniteElement {
public:
    typedef GmshMesh mesh_type;
    typedef GmshMeshSeq mesh_type_root;
    typedef GmshMesh::element_type element_type;

    // ---- declarations of the definitions cut by extract.py (signatures as in model/finiteelement.hpp) ----
    void initFETensors();
    double jacobian(std::vector<std::vector<double>> const& vertices) const;
    template <typename FEMeshType>
    double jacobian(element_type const& element, FEMeshType const& mesh) const            // finiteelement.hpp:103-105
    { return this->jacobian(mesh.vertices(element.indices)); }
    template <typename FEMeshType>
    double jacobian(element_type const& element, FEMeshType const& mesh,
                    std::vector<double> const& um, double factor = 1.) const              // finiteelement.hpp:107-110
    { return this->jacobian(mesh.vertices(element.indices, um, factor)); }
    std::vector<double> sides(element_type const& element, mesh_type const& mesh) const;
    std::vector<double> sides(element_type const& element, mesh_type const& mesh,
                              std::vector<double> const& um, double factor = 1.) const;
    template <typename FEMeshType> double measure(element_type const& element, FEMeshType const& mesh) const;
    template <typename FEMeshType> double measure(element_type const& element, FEMeshType const& mesh,
                                                  std::vector<double> const& um, double factor = 1.) const;
    std::vector<double> shapeCoeff(element_type const& element) const;
    template <typename FEMeshType> double minAngles(element_type const& element, FEMeshType const& mesh,
                                                    std::vector<double> const& um, double factor) const;
    template <typename FEMeshType> double minAngle(FEMeshType const& mesh, std::vector<double> const& um, double factor,
                                                   bool root = false) const;
    template <typename FEMeshType> bool flip(FEMeshType const& mesh, std::vector<double> const& um, double factor) const;
    void calcCohesion();
    void update(std::vector<double> const& UM_P);
    void updateSigmaDamage(double const dt);
    void updateIceDiagnostics();
    bool checkRegridding();
    void explicitSolve();
    inline void updateSigmaVP(double const e, double const Pstar, double const C, double const delta_min,
                              double const ralpha1, double const ralpha2);
    inline void updateSigmaEVP(double const dte, double const e, double const Pstar, double const C, double const delta_min);
    inline void updateSigmaMEVP(double const e, double const Pstar, double const C, double const delta_min, double const alpha);
    void updateGhosts(std::vector<double>& mesh_nodal_vec);
    // SURVEY 8(f) row 3
    inline std::pair<double, double> specificHumidity(schemes::specificHumidity scheme, int i, double temp = -999.);
    void OWBulkFluxes(std::vector<double>& Qow, std::vector<double>& Qlw, std::vector<double>& Qsw, std::vector<double>& Qlh,
                      std::vector<double>& Qsh, std::vector<double>& evap, ModelVariable& tau);
    void IABulkFluxes(const std::vector<double>& Tsurf, const std::vector<double>& snow_thick, const std::vector<double>& conc,
                      std::vector<double>& Qia, std::vector<double>& Qlw, std::vector<double>& Qsw, std::vector<double>& Qlh,
                      std::vector<double>& Qsh, std::vector<double>& I, std::vector<double>& subl, std::vector<double>& dQiadT,
                      std::vector<double>& alb_tot, ModelVariable& drag_ui, ModelVariable& drag_ti, bool bulk_for_young);
    void thermo(int dt);
    inline double windSpeedElement(const int i);
    inline double incomingLongwave(const int i);
    inline double iceOceanHeatflux(const int cpt, const double sst, const double sss, const double mld, const double dt);
    inline double freezingPoint(const double sss);
    inline std::tuple<double, double> albedo(const double Tsurf, const double hs, const double frac_pnd, const int alb_scheme,
                                             const double alb_ice, const double alb_sn, const double alb_pnd, const double I_0);
    inline void meltPonds(const int cpt, const double dt, const double hi, const double hs, const double iceSurfaceMelt,
                          const double snowMelt, const double Qia, const double rain, const double roff, const double dep2frac);
    inline void thermoWinton(const double dt, const double conc, const double voli, const double vols, const double mld,
                             const double snowfall, const double Qia, const double dQiadT, const double I, const double subl,
                             const double Tbot, double& Qio, double& hi, double& hs, double& hi_old, double& del_hi,
                             double& del_hs_mlt, double& mlt_hi_top, double& mlt_hi_bot, double& del_hi_s2i, double& Tsurf,
                             double& T1, double& T2);
    inline void thermoIce0(const double dt, const double conc, const double voli, const double vols, const double mld,
                           const double snowfall, const double Qia, const double dQiadT, const double I, const double subl,
                           const double Tbot, double& Qio, double& hi, double& hs, double& hi_old, double& del_hi,
                           double& del_hs_mlt, double& mlt_hi_top, double& mlt_hi_bot, double& del_hi_s2i, double& Tsurf);

    // diffuse() gathers to the root mesh; it returns at once for diffusivity <= 0 (FE.cpp:2762-2767), the default
    void diffuse(ModelVariable&, double diffusivity, double)
    { if (diffusivity > 0.) throw std::runtime_error("ref_fe stub: thermo.diffusivity_* > 0 is not supported"); }

    // ---- members the cut bodies read or write (names and types of model/finiteelement.hpp) ----
    OptMap vm;
    mutable Timer M_timer;
    Communicator M_comm;
    int M_rank = 0;
    mesh_type M_mesh;
    std::vector<element_type> M_elements;
    BamgMeshStub* bamgmesh = nullptr;
    int M_num_elements = 0, M_num_nodes = 0, M_local_ndof = 0;
    setup::DynamicsType M_dynamics_type = setup::DynamicsType::BBM;
    setup::BasalStressType M_basal_stress_type = setup::BasalStressType::NONE;
    setup::IceCategoryType M_ice_cat_type = setup::IceCategoryType::CLASSIC;

    double dtime_step = 0., ocean_turning_angle_rad = 0.;
    double nu0 = 0., young = 0., compaction_param = 0., undamaged_time_relaxation_sigma = 0., exponent_relaxation_sigma = 0.;
    double compression_factor = 0., exponent_compression_factor = 0., compr_strength = 0., tan_phi = 0.;
    double quad_drag_coef_water = 0., C_fix = 0., C_alea = 0., M_res_root_mesh = 0.;
    double const days_in_sec = 86400.;

    std::vector<double> M_Dunit, M_surface, M_delta_x, M_UM, M_UT, M_VT;
    std::vector<std::vector<double>> M_shape_coeff, M_B0T;
    std::vector<bool> M_mask_dirichlet;
    std::vector<int> M_neumann_nodes, M_neumann_flags;
    ExternalData M_wind, M_ocean, M_ssh, M_element_depth;

    ModelVariable M_conc, M_thick, M_snow_thick, M_damage, M_ridge_ratio, M_conc_young, M_h_young, M_hs_young;
    ModelVariable M_thick_myi, M_conc_myi, M_Cohesion, M_time_relaxation_damage, M_drag_ui, M_drag_ui_young;
    ModelVariable M_random_number, M_sst, M_sss, M_tsurf_young;
    std::vector<ModelVariable> M_sigma, M_tice, D_sigma;
    ModelVariable D_tau_a, D_tau_w, D_del_ci_ridge_myi, D_conc, D_thick, D_snow_thick, D_tsurf, D_divergence, D_dmean, D_dmax;

    // ---- thermo(): options, forcing, state and diagnostics (names and types of model/finiteelement.hpp) ----
    setup::OceanType M_ocean_type = setup::OceanType::CONSTANT;
    setup::ThermoType M_thermo_type = setup::ThermoType::WINTON;
    setup::OceanHeatfluxScheme M_Qio_type = setup::OceanHeatfluxScheme::BASIC;
    setup::FreezingPointType M_freezingpoint_type = setup::FreezingPointType::LINEAR;
    double M_current_time = 0., M_ocean_albedo = 0., M_ks = 0., M_freezingpoint_mu = 0., M_Csens_io = 0.;
    double time_relaxation_damage = 0., deltaT_relaxation_damage = 0., h_young_min = 0., h_young_max_sharp = 0., quad_drag_coef_air = 0.;
    bool M_flooding = true;
    ExternalData M_tair, M_mixrat, M_dair, M_sphuma, M_mslp, M_Qsw_in, M_Qlw_in, M_tcc, M_precip, M_snowfall, M_snowfr, M_mld,
        M_ocean_temp, M_ocean_salt;
    ModelVariable M_conc_upd, M_del_vi_tend, M_freeze_days, M_freeze_onset, M_conc_summer, M_thick_summer, M_fyi_fraction,
        M_age_det, M_age, M_pond_volume, M_lid_volume, M_drag_ti, M_drag_ti_young;
    ModelVariable D_tau_ow, D_pond_fraction, D_Qa, D_Qsw, D_Qlw, D_Qsh, D_Qlh, D_Qo, D_Qnosun, D_Qsw_ocean, D_Qassim, D_delS,
        D_fwflux_ice, D_fwflux, D_brine, D_evap, D_rain, D_vice_melt, D_del_vi_young, D_del_hi, D_del_hi_young, D_newice,
        D_mlt_top, D_mlt_bot, D_snow2ice, D_albedo, D_sialb, D_del_ci_mlt_myi, D_del_vi_mlt_myi, D_del_ci_rplnt_myi,
        D_del_vi_rplnt_myi;

    // ghost exchange lists (FE.cpp:14003-14088 fills them; here the harness does)
    std::vector<std::vector<int>> M_extract_local_index, M_local_ghosts_local_index;
    std::vector<int> M_recipients_proc_id, M_local_ghosts_proc_id;
};

}  // namespace Nextsim

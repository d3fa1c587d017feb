// =====================================================================================
// oracle/nextsim_oracle.cpp  --  TEST INFRASTRUCTURE ONLY (not product code)
//
// CPU restatement (C++17, FP64, single thread per rank) of the neXtSIM explicit
// momentum / rheology hot path.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library; the product
// path (libnsx.so) never links, imports or calls it.
//
// PARITY PINNED against the reference's own text (round 2): oracle/ref_fe cuts the definitions of
// explicitSolve, update, updateSigmaDamage, updateSigmaVP/EVP/MEVP, updateGhosts, sides, measure, shapeCoeff,
// jacobian, minAngle, flip, checkRegridding, updateIceDiagnostics, calcCohesion and initFETensors out of
// /root/reference/model/finiteelement.cpp at build time, compiles that text against a stub class
// (oracle/ref_fe/stub_fe.hpp; the real header needs Boost / MPI / Gmsh / NetCDF, absent here) into
// oracle/_ref/libref_fe.so, and tests/test_ref_fe_cpu.py requires this restatement to reproduce those
// bodies BIT FOR BIT (np.array_equal on every output) for BBM / EVP / mEVP, 1-4 ranks, young ice, open
// boundaries, Lemieux basal stress.  The reference ships no test, fixture or stored output for this path
//  and its executable cannot be built in this image, so that comparison is the pin.
// NOT pinned (no reference code can run it here): nodalGrid() -- it needs Gmsh-written partition tags and
// boost::mpi; the restatement follows gmshmesh.cpp:856-1498 line by line.
// The restatement is built with -O2 -ffp-contract=off, like libref_fe.so; the golden vectors under tests/golden/
// are generated from it (tests/golden/make_golden.py).
// PINNED against the reference's own code: bamgTables() -- contrib/bamg compiles standalone
// and is built UNMODIFIED from /root/reference into oracle/_ref/libref_bamg.so
// (oracle/ref_bamg/Makefile); tests/test_ref_bamg_cpu.py checks NodalElementConnectivity and
// NodalConnectivity of BamgConvertMeshx (the call of FE.cpp:77-80) against bamgTables()
// bit for bit on root and partition-local meshes.
// Also built unmodified: contrib/mapx (oracle/ref_mapx) for GmshMesh::lat(), which pins the product's
// nsx_mapx_latlon (tests/test_ref_mapx_cpu.py); the oracle itself takes lat as an input.
//
// Reference files followed (paths relative to /root/reference):
//   model/finiteelement.cpp
//     1491-1507   initFETensors          -> Rank::initFETensors
//     1613-1618   jacobian               -> jacobian()
//     1642-1663   sides(um)              -> sides()
//     1929-1933   measure(um)            -> measure()
//     1951-1964   shapeCoeff             -> shapeCoeff()
//     150-271     bcMarkedNodes          -> orc_bc_marked_nodes
//     3909-3914   calcCohesion           -> orc_calc_cohesion (+ 11459-11475 random field)
//     3919-4132   update                 -> update()
//     4137-4260   updateSigmaDamage      -> updateSigmaDamage()
//     10182-10643 explicitSolve          -> explicitSolve() (lock-step over ranks)
//     10649-10726 updateSigmaVP/EVP/MEVP -> updateSigmaVP()
//     13963-13996 updateGhosts           -> updateGhosts()
//     14003-14105 initUpdateGhosts, globalNumToprocId -> initUpdateGhosts()
//   core/src/gmshmesh.cpp 856-1498 nodalGrid, core/include/entities.hpp 105-134
//                                        -> orc_nodal_grid
//   contrib/bamg/src/Mesh.cpp 514-543, 583-629, 798-865 (connectivity tables)
//                                        -> orc_bamg_tables
//   model/constants.hpp 56-86            -> namespace physical
//   SURVEY.md section 8(f) rows 1-2 (the callers either side of the path):
//     model/finiteelement.cpp 1758-1768 minAngles(um), 1795-1816 minAngle(um), 1824-1839 flip,
//       8298-8309 checkRegridding (local part)   -> orc_check_regridding
//     model/finiteelement.cpp 7860-7900 updateIceDiagnostics -> orc_update_ice_diagnostics
//     model/externaldata.cpp 366-436 ExternalData::get / 437-455 getVector
//                                        -> orc_external_data_get_vector
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <numeric>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace physical {           // model/constants.hpp:56-86
const double rhoi = 917.;
const double rhow = 1025.;
const double rhos = 330.;
const double gravity = 9.80616;
const double omega = 7.292e-5;
const double rhoa = 1.22;
}
static const double PI = 3.14159265358979323846;   // core: #define PI M_PI
static const double days_in_sec = 86400.;

extern "C" {
typedef struct OrcParams {
    int dynamics_type;        // setup::DynamicsType  BBM=0, EVP=3, mEVP=4 (enums.hpp:142-149)
    int basal_stress_type;    // NONE=0, LEMIEUX=1
    int ice_cat_type;         // CLASSIC=0, YOUNG_ICE=1
    int substeps;             // dynamics.substeps
    int equal_ridging;        // age.equal_ridging
    int newice_type;          // thermo.newice_type
    int use_young_ice_in_myi_reset; // age.include_young_ice
    int stop_after_substeps;  // debug: run only the first k sub-cycles (0 = all)
    int skip_ow_smoother;     // debug: skip FE.cpp:10578-10611
    int pad_;
    double dtime_step;
    double ocean_turning_angle_rad;
    double min_h, min_c;
    double young, nu0, tan_phi, compr_strength, compaction_param;
    double undamaged_time_relaxation_sigma, exponent_relaxation_sigma;
    double compression_factor, exponent_compression_factor;
    double quad_drag_coef_water;
    double evp_e, evp_Pstar, evp_C, evp_dmin, mevp_alpha, mevp_beta;
    double basal_k1, basal_k2, basal_Cb, basal_u0;
} OrcParams;
}

namespace {

struct Rank
{
    int rank = 0, nranks = 1;
    std::map<std::string, std::vector<double>> d;
    std::map<std::string, std::vector<int>> i;

    // sizes (FE.cpp:80-91)
    int M_num_nodes = 0, M_local_ndof = 0, M_num_elements = 0, M_local_nelements = 0;
    int nec_width = 0, nc_width = 0;   // bamg table widths

    // halo lists (FE.cpp:14003-14088)
    std::vector<std::vector<int>> M_extract_local_index, M_local_ghosts_local_index;
    std::vector<int> M_recipients_proc_id, M_local_ghosts_proc_id;

    // scratch of explicitSolve shared across phases
    std::vector<double> element_mass, rlmass_matrix, node_mass, C_bu, grad_ssh, fcor, VTM, M_Dunit;

    std::vector<double>& D(const char* n) { return d[n]; }
    std::vector<int>& I(const char* n) { return i[n]; }
};

// ---- geometry helpers ---------------------------------------------------------------
// GmshMesh::vertices(indices, um, factor)  core/src/gmshmesh.cpp:1929-1939
static inline void vertices(Rank& R, int cpt, double v[3][2])
{
    auto const& idx = R.I("indices");
    auto const& X = R.D("coordX");
    auto const& Y = R.D("coordY");
    auto const& um = R.D("M_UM");
    for (int i=0; i<3; ++i)
    {
        int const n = idx[3*cpt+i];        // 1-based
        v[i][0] = X[n-1];
        v[i][1] = Y[n-1];
        for (int k=0; k<2; ++k)
            v[i][k] += 1.*um[n-1+k*R.M_num_nodes];
    }
}

// FE.cpp:1613-1618
static inline double jacobian(double const v[3][2])
{
    double jac = (v[1][0]-v[0][0])*(v[2][1]-v[0][1]);
    jac -= (v[2][0]-v[0][0])*(v[1][1]-v[0][1]);
    return jac;
}

// FE.cpp:1929-1933
static inline double measure(Rank& R, int cpt)
{
    double v[3][2];
    vertices(R, cpt, v);
    return (1./2)*std::abs(jacobian(v));
}

// FE.cpp:1642-1663
static inline void sides(Rank& R, int cpt, double side[3])
{
    double v[3][2];
    vertices(R, cpt, v);
    side[0] = std::hypot(v[1][0]-v[0][0], v[1][1]-v[0][1]);
    side[1] = std::hypot(v[2][0]-v[1][0], v[2][1]-v[1][1]);
    side[2] = std::hypot(v[2][0]-v[0][0], v[2][1]-v[0][1]);
}

// FE.cpp:1951-1964
static inline void shapeCoeff(Rank& R, int cpt, double coeff[6])
{
    double v[3][2];
    vertices(R, cpt, v);
    double const jac = jacobian(v);
    for (int k=0; k<3; ++k)
    {
        int const kp1 = (k+1)%3;
        int const kp2 = (k+2)%3;
        coeff[k]   = (v[kp1][1]-v[kp2][1])/jac;
        coeff[k+3] = (v[kp2][0]-v[kp1][0])/jac;
    }
}

// FE.cpp:1491-1507
static void initFETensors(Rank& R, OrcParams const& P)
{
    R.M_Dunit.assign(9,0);
    double const Dunit_factor=1./(1.-P.nu0*P.nu0);
    R.M_Dunit[0]= Dunit_factor * 1.;
    R.M_Dunit[1]= Dunit_factor * P.nu0;
    R.M_Dunit[3]= Dunit_factor * P.nu0;
    R.M_Dunit[4]= Dunit_factor * 1.;
    R.M_Dunit[8]= Dunit_factor * (1.-P.nu0)/2.;
}

// ---- BBM  FE.cpp:4137-4260 -------------------------------------------------------------
static void updateSigmaDamage(Rank& R, OrcParams const& P, double const dt)
{
    double const sqrt_nu_rhoi = std::sqrt( 2.*(1.+P.nu0)*physical::rhoi );
    const double min_c = 0.1;                       // Q4: hard-coded (FE.cpp:4146)

    auto& M_conc = R.D("M_conc");
    auto& M_thick = R.D("M_thick");
    auto& M_damage = R.D("M_damage");
    auto& M_VT = R.D("M_VT");
    auto& M_B0T = R.D("M_B0T");
    auto& M_Cohesion = R.D("M_Cohesion");
    auto& M_delta_x = R.D("M_delta_x");
    auto& M_trd = R.D("M_time_relaxation_damage");
    auto const& idx = R.I("indices");
    std::vector<double>* M_sigma[3] = { &R.D("M_sigma0"), &R.D("M_sigma1"), &R.D("M_sigma2") };
    int const M_num_nodes = R.M_num_nodes;

    for (int cpt=0; cpt < R.M_num_elements; ++cpt)
    {
        if ( M_conc[cpt] <= min_c )
        {
            M_damage[cpt] = 0.;
            for(int i=0;i<3;i++)
                (*M_sigma[i])[cpt] = 0.;
            continue;
        }

        double epsilon_veloc[3] = {0.,0.,0.};
        for(int i=0;i<3;i++)
        {
            for(int j=0;j<3;j++)
            {
                epsilon_veloc[i] += M_B0T[18*cpt + i*6 + 2*j]*M_VT[idx[3*cpt+j]-1];
                epsilon_veloc[i] += M_B0T[18*cpt + i*6 + 2*j + 1]*M_VT[idx[3*cpt+j]-1+M_num_nodes];
            }
        }

        double sigma_n = ((*M_sigma[0])[cpt]+(*M_sigma[1])[cpt])*0.5;
        double const expC = std::exp(P.compaction_param*(1.-M_conc[cpt]));
        double const time_viscous = P.undamaged_time_relaxation_sigma*std::pow((1.-M_damage[cpt])*expC,P.exponent_relaxation_sigma-1.);

        double tildeP;
        if ( sigma_n < 0. )
        {
            double const Pmax = std::pow(M_thick[cpt], P.exponent_compression_factor)*P.compression_factor*expC;
            tildeP = std::min(1., -Pmax/sigma_n);
        } else {
            tildeP = 0.;
        }

        double const multiplicator = std::min( 1. - 1e-12,
                time_viscous/(time_viscous+dt*(1.-tildeP)) );

        double const elasticity = P.young*(1.-M_damage[cpt])*expC;

        for(int i=0;i<3;i++)
        {
            for(int j=0;j<3;j++)
                (*M_sigma[i])[cpt] += dt*elasticity*R.M_Dunit[3*i + j]*epsilon_veloc[j];

            (*M_sigma[i])[cpt] *= multiplicator;
        }

        double const sigma_s = std::hypot(((*M_sigma[0])[cpt]-(*M_sigma[1])[cpt])/2.,(*M_sigma[2])[cpt]);
        sigma_n = ((*M_sigma[0])[cpt]+(*M_sigma[1])[cpt])*0.5;

        double dcrit;
        if ( sigma_n < -P.compr_strength )
            dcrit = -P.compr_strength/sigma_n;
        else
            dcrit = M_Cohesion[cpt]/(sigma_s+P.tan_phi*sigma_n);

        if ( (0.<dcrit) && (dcrit<1.) )
        {
            double const rtd = std::sqrt(elasticity)/(M_delta_x[cpt]*sqrt_nu_rhoi);
            double const del_damage = (1.0-M_damage[cpt])*(1.0-dcrit)*dt*rtd;
            M_damage[cpt] += del_damage;

            for (int i=0;i<3;i++)
                (*M_sigma[i])[cpt] -= (*M_sigma[i])[cpt]*(1.-dcrit)*dt*rtd;
        }

        M_damage[cpt] = std::max( 0., M_damage[cpt]
                - dt/M_trd[cpt]*std::exp(P.compaction_param*(1.-M_conc[cpt])) );
    }
}

// ---- EVP / mEVP  FE.cpp:10649-10726 ------------------------------------------------------
static void updateSigmaVP(Rank& R, double const e, double const Pstar, double const C,
                          double const delta_min, double const ralpha1, double const ralpha2)
{
    double const re2 = 1./(e*e);
    auto& M_conc = R.D("M_conc");
    auto& M_thick = R.D("M_thick");
    auto& M_VT = R.D("M_VT");
    auto& M_shape_coeff = R.D("M_shape_coeff");
    auto const& idx = R.I("indices");
    auto& s0 = R.D("M_sigma0"); auto& s1 = R.D("M_sigma1"); auto& s2 = R.D("M_sigma2");
    int const M_num_nodes = R.M_num_nodes;

    for ( int cpt=0; cpt<R.M_num_elements; cpt++ )
    {
        if ( M_thick[cpt] == 0. )
        {
            s0[cpt] = 0.; s1[cpt] = 0.; s2[cpt] = 0.;
            continue;
        }

        double eps11 = 0.;
        double eps22 = 0.;
        double eps12 = 0.;
        for(int i=0; i<3; i++)
        {
            double const u = M_VT[idx[3*cpt+i]-1];
            double const v = M_VT[idx[3*cpt+i]-1 + M_num_nodes];
            double const dxN = M_shape_coeff[6*cpt+i];
            double const dyN = M_shape_coeff[6*cpt+i+3];
            eps11 += dxN*u;
            eps22 += dyN*v;
            eps12 += 0.5*( dxN*v + dyN*u );
        }

        double const eps1 = eps11 + eps22;
        double const eps2 = eps11 - eps22;

        double const delta = std::sqrt( eps1*eps1 + (eps2*eps2 + 4*eps12*eps12)*re2 );
        double const Pp = Pstar*std::exp(-C*(1.-M_conc[cpt]));
        double const zeta = Pp / ( delta + delta_min );

        double sigma1 = s0[cpt] + s1[cpt];
        double sigma2 = s0[cpt] - s1[cpt];

        sigma1 += ralpha1*( zeta*(eps1-delta) - sigma1 );
        sigma2 += ralpha2*( zeta*eps2*re2 - sigma2 );
        s2[cpt] += ralpha2*( zeta*eps12*re2 - s2[cpt] );

        s0[cpt] = 0.5*(sigma1 + sigma2);
        s1[cpt] = 0.5*(sigma1 - sigma2);
    }
}

// ---- updateGhosts  FE.cpp:13963-13996 (in-process replay of the MPI p2p) ------------------
static void updateGhosts(std::vector<Rank*>& W, const char* name)
{
    int const n = (int)W.size();
    if (n == 1) return;       // a single rank has no recipients (lists are empty)
    // pack on every rank ("send")
    std::vector<std::vector<std::vector<double>>> msg(n, std::vector<std::vector<double>>(n));
    for (int r=0; r<n; ++r)
    {
        Rank& R = *W[r];
        auto& vec = R.D(name);
        for (int i=0; i<(int)R.M_extract_local_index.size(); i++)
        {
            int const srl = (int)R.M_extract_local_index[i].size();
            msg[r][i].resize(2*srl);
            for (int j=0; j<srl; j++)
            {
                msg[r][i][j] = vec[R.M_extract_local_index[i][j]];
                msg[r][i][j+srl] = vec[R.M_extract_local_index[i][j]+R.M_num_nodes];
            }
        }
    }
    // unpack on every rank ("recv" from proc, tag proc)
    for (int r=0; r<n; ++r)
    {
        Rank& R = *W[r];
        auto& vec = R.D(name);
        for (int i=0; i<(int)R.M_local_ghosts_local_index.size(); i++)
        {
            int const srl = (int)R.M_local_ghosts_local_index[i].size();
            if (srl == 0) continue;
            auto const& ghost_update_values = msg[i][r];
            if ((int)ghost_update_values.size() != 2*srl)
                throw std::logic_error("updateGhosts: message size mismatch");
            for (int j=0; j<srl; j++)
            {
                vec[R.M_local_ghosts_local_index[i][j]] = ghost_update_values[j];
                vec[R.M_local_ghosts_local_index[i][j]+R.M_num_nodes] = ghost_update_values[j+srl];
            }
        }
    }
}

// ---- explicitSolve  FE.cpp:10182-10643 ------------------------------------------------------
// Phase A: everything before the sub-cycle loop (10184-10416), one rank.
static void explicitSolve_prep(Rank& R, OrcParams const& P)
{
    int const M_num_elements = R.M_num_elements;
    int const M_num_nodes = R.M_num_nodes;
    auto const& idx = R.I("indices");
    auto const& ghostNodes = R.I("ghostNodes");
    auto const& M_mask_dirichlet = R.I("M_mask_dirichlet");
    auto& ssh = R.D("M_ssh");                     // 10219
    auto& M_conc = R.D("M_conc");
    auto& M_thick = R.D("M_thick");
    auto& M_snow_thick = R.D("M_snow_thick");
    auto& M_conc_young = R.D("M_conc_young");
    auto& M_h_young = R.D("M_h_young");
    auto& M_hs_young = R.D("M_hs_young");
    auto& M_element_depth = R.D("M_element_depth");
    auto& M_VT = R.D("M_VT");
    auto& M_wind = R.D("M_wind");
    auto& M_drag_ui = R.D("M_drag_ui");
    auto& M_drag_ui_young = R.D("M_drag_ui_young");
    auto& lat = R.D("lat");

    double const k1 = P.basal_k1, k2 = P.basal_k2, Cb = P.basal_Cb;

    auto& M_delta_x = R.D("M_delta_x");         M_delta_x.resize(M_num_elements);
    auto& M_surface = R.D("M_surface");         M_surface.resize(M_num_elements);
    auto& M_shape_coeff = R.D("M_shape_coeff"); M_shape_coeff.resize(6*(size_t)M_num_elements);
    auto& M_B0T = R.D("M_B0T");                 M_B0T.resize(18*(size_t)M_num_elements);
    auto& D_tau_a = R.D("D_tau_a");             D_tau_a.resize(2*(size_t)M_num_nodes);
    auto& D_tau_w = R.D("D_tau_w");             D_tau_w.resize(2*(size_t)M_num_nodes);

    auto& element_mass = R.element_mass;   element_mass.assign(M_num_elements, 0.);
    auto& rlmass_matrix = R.rlmass_matrix; rlmass_matrix.assign(M_num_nodes, 0.);
    auto& node_mass = R.node_mass;         node_mass.assign(M_num_nodes, 0.);
    auto& C_bu = R.C_bu;                   C_bu.assign(M_num_nodes, 0.);
    auto& grad_ssh = R.grad_ssh;           grad_ssh.assign(2*(size_t)M_num_nodes, 0.);

    for ( int cpt=0; cpt<M_num_elements; ++cpt )
    {
        double my_sides[3];
        sides(R, cpt, my_sides);
        // Q1: std::accumulate(begin,end,0) -> *int* accumulator, then integer division by size_t 3
        int acc = 0;
        for (int s=0; s<3; ++s)
            acc = acc + my_sides[s];            // int = (double)(int + double), truncation each step
        M_delta_x[cpt] = acc/(size_t)3;
        M_surface[cpt] = measure(R, cpt);
        double shapecoeff[6];
        shapeCoeff(R, cpt, shapecoeff);
        double* B0T = &M_B0T[18*(size_t)cpt];
        for (int k=0; k<18; ++k) B0T[k] = 0;
        for (int i=0; i<3; ++i)
        {
            B0T[2*i] = shapecoeff[i];
            B0T[2*i+13] = shapecoeff[i];
            B0T[2*i+7] = shapecoeff[i+3];
            B0T[2*i+12] = shapecoeff[i+3];
        }
        for (int k=0; k<6; ++k) M_shape_coeff[6*(size_t)cpt+k] = shapecoeff[k];

        double total_concentration=M_conc[cpt];
        double total_thickness=M_thick[cpt];
        double total_snow=M_snow_thick[cpt];

        if(P.ice_cat_type==1)
        {
            total_concentration += M_conc_young[cpt];
            total_thickness     += M_h_young[cpt];
            total_snow          += M_hs_young[cpt];
        }

        if ( total_concentration > 0. )
            element_mass[cpt] = (physical::rhoi*total_thickness + physical::rhos*total_snow)/total_concentration;
        else
            element_mass[cpt] = 0.;

        double element_ssh = 0;
        for (int i=0; i<3; ++i)
            element_ssh += ssh[idx[3*cpt+i]-1];
        element_ssh /= 3.;

        double max_keel_depth=28;
        double mean_keel_depth;
        double critical_h = 0.;
        double critical_h_mod = 0.;
        double const min_water_depth = 2.;
        double const depth_eff = std::max(0., element_ssh
                + std::max(min_water_depth, M_element_depth[cpt]));
        double const g3rd = physical::gravity/3.;
        switch ( P.basal_stress_type )
        {
            case 0:
                critical_h     = 0.;
                critical_h_mod = 0.;
                break;
            case 1:
                mean_keel_depth = k1 * M_thick[cpt];
                mean_keel_depth = std::min( mean_keel_depth, M_conc[cpt] * max_keel_depth );
                critical_h     = M_conc[cpt] * depth_eff / k1;
                critical_h_mod = mean_keel_depth / k1;
                break;
        }

        double const element_C_bu = k2*std::max(0., critical_h_mod-critical_h)*std::exp(-Cb*(1.-M_conc[cpt]));
        for (int i=0; i<3; ++i)
        {
            int const idx_node = idx[3*cpt+i]-1;
            rlmass_matrix[idx_node] += M_surface[cpt];
            node_mass[idx_node] += element_mass[cpt]*M_surface[cpt];
            C_bu[idx_node]  = std::max(C_bu[idx_node], element_C_bu);
        }

        double const m_g_A3rd = element_mass[cpt]*M_surface[cpt]*g3rd;
        double const* dxN = &M_shape_coeff[6*(size_t)cpt];
        for (int i=0; i<3; ++i)
        {
            int const i_indx = idx[3*cpt+i]-1;

            // NB: node_mass is the *running* sum at this point of the element loop (10328)
            if ( M_mask_dirichlet[i_indx] || node_mass[i_indx]==0. || ghostNodes[3*cpt+i] )
                continue;

            int const u_indx = i_indx;
            int const v_indx = i_indx + M_num_nodes;

            for ( int j=0; j<3; ++j )
            {
                int const j_indx = idx[3*cpt+j]-1;
                grad_ssh[u_indx] -= dxN[j] * m_g_A3rd * ssh[j_indx];
                grad_ssh[v_indx] -= dxN[j+3] * m_g_A3rd * ssh[j_indx];
            }
        }
    }

    // prep nodes 10356-10416
    auto& fcor = R.fcor; fcor.assign(M_num_nodes, 0.);
    auto& VTM = R.VTM;   VTM.assign(2*(size_t)M_num_nodes, 0.);
    auto const& NEC = R.D("NodalElementConnectivity");
    for ( int i=0; i<M_num_nodes; ++i )
    {
        const int u_indx = i;
        const int v_indx = i+M_num_nodes;

        if ( node_mass[i]==0. )
        {
            M_VT[u_indx] = 0.;
            M_VT[v_indx] = 0.;
        }

        double drag = 0.;
        double surface = 0;
        int num_elements = R.nec_width;
        for (int j=0; j<num_elements; j++)
        {
            // Q6: (int)NaN - 1 is negative on x86 (cvttsd2si -> INT_MIN); restated explicitly
            double const raw = NEC[(size_t)num_elements*i+j];
            if (std::isnan(raw)) continue;
            int elt_num = (int)raw-1;
            if ( elt_num < 0 ) continue;

            double dragp = M_drag_ui[elt_num];
            if ( P.ice_cat_type==1 && M_conc[elt_num]+M_conc_young[elt_num] > 0. )
                dragp = (M_drag_ui[elt_num]*M_conc[elt_num]+M_drag_ui_young[elt_num]*M_conc_young[elt_num])
                    /(M_conc[elt_num]+M_conc_young[elt_num]);

            drag += dragp * M_surface[elt_num];
            surface += M_surface[elt_num];
        }
        drag *= physical::rhoa * std::hypot(M_wind[u_indx],M_wind[v_indx]) / surface;

        D_tau_a[u_indx] = drag * M_wind[u_indx];
        D_tau_a[v_indx] = drag * M_wind[v_indx];

        fcor[i] = 2*physical::omega*std::sin(lat[i]*PI/180.);

        rlmass_matrix[i] = 1./rlmass_matrix[i];
        node_mass[i] *= rlmass_matrix[i];
        rlmass_matrix[i] *= 3.;

        VTM[u_indx] = M_VT[u_indx];
        VTM[v_indx] = M_VT[v_indx];
    }
}

// Phase B: one sub-cycle up to (not including) updateGhosts, one rank (10425-10530)
static void explicitSolve_substep(Rank& R, OrcParams const& P, double const dte)
{
    int const M_num_nodes = R.M_num_nodes;
    auto const& idx = R.I("indices");
    auto const& ghostNodes = R.I("ghostNodes");
    auto const& M_mask_dirichlet = R.I("M_mask_dirichlet");
    auto& M_thick = R.D("M_thick");
    auto& M_surface = R.D("M_surface");
    auto& M_shape_coeff = R.D("M_shape_coeff");
    auto& M_VT = R.D("M_VT");
    auto& M_ocean = R.D("M_ocean");
    auto& D_tau_a = R.D("D_tau_a");
    auto& lat = R.D("lat");
    auto& s0 = R.D("M_sigma0"); auto& s1 = R.D("M_sigma1"); auto& s2 = R.D("M_sigma2");
    auto const& tau_wi = R.D("tau_wi");           // optional OASIS slot (empty -> unused)
    bool const have_tau_wi = (tau_wi.size() == 2*(size_t)M_num_nodes);

    double const cos_ocean_turning_angle = std::cos(P.ocean_turning_angle_rad);
    double const sin_ocean_turning_angle = std::sin(P.ocean_turning_angle_rad);
    double const min_m = physical::rhoi*P.min_h;
    double const u0 = P.basal_u0;

    switch(P.dynamics_type)
    {
        case 3: {   // EVP 10705-10715
            double const T = P.dtime_step / 3.;
            double const ralpha1 = 0.5*dte/T;
            double const ralpha2 = 0.5*dte/T*P.evp_e*P.evp_e;
            updateSigmaVP(R, P.evp_e, P.evp_Pstar, P.evp_C, P.evp_dmin, ralpha1, ralpha2);
            break; }
        case 4:     // mEVP 10721-10726
            updateSigmaVP(R, P.evp_e, P.evp_Pstar, P.evp_C, P.evp_dmin, 1./P.mevp_alpha, 1./P.mevp_alpha);
            break;
        case 0:
            updateSigmaDamage(R, P, dte);
            break;
    }

    std::vector<double> grad_terms = R.grad_ssh;
    for ( int cpt=0; cpt<R.M_num_elements; ++cpt )
    {
        double const* dxN = &M_shape_coeff[6*(size_t)cpt];
        double const volume = M_thick[cpt]*M_surface[cpt];
        for (int i=0; i<3; ++i)
        {
            int const i_indx = idx[3*cpt+i]-1;
            if ( M_mask_dirichlet[i_indx] || R.node_mass[i_indx]==0. || ghostNodes[3*cpt+i] )
                continue;

            int const u_indx = i_indx;
            int const v_indx = i_indx + M_num_nodes;

            grad_terms[u_indx] -= volume*( s0[cpt]*dxN[i] + s2[cpt]*dxN[i+3] );
            grad_terms[v_indx] -= volume*( s2[cpt]*dxN[i] + s1[cpt]*dxN[i+3] );
        }
    }

    for ( int i=0; i<R.M_local_ndof; ++i )
    {
        if ( M_mask_dirichlet[i] || R.node_mass[i]==0. )
            continue;

        int u_indx = i;
        int v_indx = i+M_num_nodes;

        double dtep, delu, delv;
        if ( P.dynamics_type == 4 )
        {
            double const b_mevp = P.mevp_beta + 1.;
            delu = (R.VTM[u_indx]-M_VT[u_indx])/b_mevp;
            delv = (R.VTM[v_indx]-M_VT[v_indx])/b_mevp;
            dtep = dte/b_mevp;
        } else {
            delu = 0.;
            delv = 0.;
            dtep = dte;
        }

        double const dte_over_mass = dtep/std::max(min_m, R.node_mass[i]);
        double const uice = M_VT[u_indx];
        double const vice = M_VT[v_indx];

        double const c_prime = physical::rhow*P.quad_drag_coef_water*std::hypot(M_ocean[u_indx]-uice, M_ocean[v_indx]-vice);

        double const tau_b = R.C_bu[i]/(std::hypot(uice,vice)+u0);
        double const alpha  = 1. + dte_over_mass*( c_prime*cos_ocean_turning_angle + tau_b );
        double const beta   = dtep*R.fcor[i] + dte_over_mass*c_prime*std::copysign(sin_ocean_turning_angle, lat[i]);
        double const rdenom = 1./( alpha*alpha + beta*beta );

        double tau_x = D_tau_a[u_indx];
        if (have_tau_wi) tau_x = tau_x + tau_wi[u_indx];
        tau_x = tau_x + c_prime*( M_ocean[u_indx]*cos_ocean_turning_angle - M_ocean[v_indx]*std::copysign(sin_ocean_turning_angle, lat[i]) );
        double tau_y = D_tau_a[v_indx];
        if (have_tau_wi) tau_y = tau_y + tau_wi[v_indx];
        tau_y = tau_y + c_prime*( M_ocean[v_indx]*cos_ocean_turning_angle + M_ocean[u_indx]*std::copysign(sin_ocean_turning_angle, lat[i]) );

        double const grad_x = grad_terms[u_indx]*R.rlmass_matrix[i];
        double const grad_y = grad_terms[v_indx]*R.rlmass_matrix[i];

        M_VT[u_indx]  = alpha*uice + beta*vice + dte_over_mass*( alpha*(grad_x + tau_x) + beta*(grad_y + tau_y) ) + alpha*delu + beta*delv;
        M_VT[u_indx] *= rdenom;

        M_VT[v_indx]  = alpha*vice - beta*uice + dte_over_mass*( alpha*(grad_y + tau_y) - beta*(grad_x + tau_x) ) + alpha*delv - beta*delu;
        M_VT[v_indx] *= rdenom;
    }
}

// move mesh 10539-10553 / 10559-10573
static void moveMesh(Rank& R, double const dt_move)
{
    auto& M_UM = R.D("M_UM");
    auto& M_UT = R.D("M_UT");
    auto& M_VT = R.D("M_VT");
    auto const& M_neumann_nodes = R.I("M_neumann_nodes");
    std::vector<double> UM_P = M_UM;
    for (size_t nd=0; nd<M_UM.size(); ++nd)
    {
        M_UM[nd] += dt_move*M_VT[nd];
        M_UT[nd] += dt_move*M_VT[nd];
    }
    for (const int& nd : M_neumann_nodes)
        M_UM[nd] = UM_P[nd];
}

// one OW smoother sweep on one rank (10582-10608)
static void owSmootherSweep(Rank& R)
{
    int const M_num_nodes = R.M_num_nodes;
    auto& M_VT = R.D("M_VT");
    auto const& M_mask_dirichlet = R.I("M_mask_dirichlet");
    auto const& NC = R.D("NodalConnectivity");
    int const max_num_neighbours = R.nc_width;
    std::vector<double> const u = M_VT;
    for ( int i=0; i<R.M_local_ndof; ++i )
    {
        int const u_indx = i;
        int const v_indx = i+M_num_nodes;

        if ( M_mask_dirichlet[i] || R.node_mass[i]!=0. )
            continue;

        M_VT[u_indx] = 0.;
        M_VT[v_indx] = 0.;

        int num_neighbours = (int)NC[(size_t)max_num_neighbours*(i+1) - 1];
        for ( int j=0; j<num_neighbours; ++j )
        {
            int const nni = (int)NC[(size_t)max_num_neighbours*i + j] - 1;
            M_VT[u_indx] += u[nni];
            M_VT[v_indx] += u[nni + M_num_nodes];
        }
        M_VT[u_indx] /= num_neighbours;
        M_VT[v_indx] /= num_neighbours;
    }
}

// 10613-10640
static void tauwAndOWMove(Rank& R, OrcParams const& P)
{
    int const M_num_nodes = R.M_num_nodes;
    auto& M_UM = R.D("M_UM");
    auto& M_UT = R.D("M_UT");
    auto& M_VT = R.D("M_VT");
    auto& M_ocean = R.D("M_ocean");
    auto& D_tau_w = R.D("D_tau_w");
    auto const& M_mask_dirichlet = R.I("M_mask_dirichlet");
    auto const& M_neumann_nodes = R.I("M_neumann_nodes");
    double const dtime_step = P.dtime_step;

    std::vector<double> UM_P = M_UM;
    for ( int i=0; i<M_num_nodes; ++i )
    {
        int const u_indx = i;
        int const v_indx = i+M_num_nodes;

        double const uice = 0.5*(M_VT[u_indx] + R.VTM[u_indx]);
        double const vice = 0.5*(M_VT[v_indx] + R.VTM[v_indx]);
        double const c_prime = physical::rhow*P.quad_drag_coef_water*std::hypot(M_ocean[u_indx]-uice, M_ocean[v_indx]-vice);
        D_tau_w[u_indx] = c_prime*( uice - M_ocean[u_indx] );
        D_tau_w[v_indx] = c_prime*( vice - M_ocean[v_indx] );

        if ( M_mask_dirichlet[i] || R.node_mass[i]!=0. )
            continue;

        M_UM[u_indx] += dtime_step*M_VT[u_indx];
        M_UM[v_indx] += dtime_step*M_VT[v_indx];

        M_UT[u_indx] += dtime_step*M_VT[u_indx];
        M_UT[v_indx] += dtime_step*M_VT[v_indx];
    }

    for (const int& nd : M_neumann_nodes)
        M_UM[nd] = UM_P[nd];
}

// Lock-step replay of explicitSolve over all ranks of the communicator; ranks interact only
// inside updateGhosts, so phase-by-phase execution is equivalent to the SPMD original.
static void explicitSolve(std::vector<Rank*>& W, OrcParams const& P)
{
    int const steps = P.substeps;
    double const dte = P.dtime_step/double(steps);
    for (Rank* R : W) { initFETensors(*R, P); explicitSolve_prep(*R, P); }

    int const nrun = (P.stop_after_substeps > 0) ? std::min(steps, P.stop_after_substeps) : steps;
    for ( int s=0; s<nrun; s++ )
    {
        for (Rank* R : W) explicitSolve_substep(*R, P, dte);
        updateGhosts(W, "M_VT");
        if ( P.dynamics_type != 4 )
            for (Rank* R : W) moveMesh(*R, dte);
    }

    if ( P.dynamics_type == 4 )
        for (Rank* R : W) moveMesh(*R, P.dtime_step);

    if (!P.skip_ow_smoother)
    {
        for ( int nit=0; nit<50; ++nit )
        {
            for (Rank* R : W) owSmootherSweep(*R);
            updateGhosts(W, "M_VT");
        }
    }

    for (Rank* R : W) tauwAndOWMove(*R, P);
}

// ---- update  FE.cpp:3919-4132 (diffuse() is a no-op at default diffusivity, 2762-2767) ------
static void update(Rank& R, OrcParams const& P)
{
    bool equal_ridging = P.equal_ridging;
    int const newice_type = P.newice_type;
    bool const use_young_ice_in_myi_reset = P.use_young_ice_in_myi_reset;
    bool const young = (P.ice_cat_type==1);

    auto const& idx = R.I("indices");
    auto const& M_neumann_flags = R.I("M_neumann_flags");
    auto& M_surface = R.D("M_surface");
    auto& M_conc = R.D("M_conc");
    auto& M_thick = R.D("M_thick");
    auto& M_snow_thick = R.D("M_snow_thick");
    auto& M_thick_myi = R.D("M_thick_myi");
    auto& M_conc_myi = R.D("M_conc_myi");
    auto& M_ridge_ratio = R.D("M_ridge_ratio");
    auto& M_h_young = R.D("M_h_young");
    auto& M_conc_young = R.D("M_conc_young");
    auto& M_hs_young = R.D("M_hs_young");
    auto& D_del_ci_ridge_myi = R.D("D_del_ci_ridge_myi");
    D_del_ci_ridge_myi.resize(R.M_num_elements);
    std::vector<double>* M_sigma[3] = { &R.D("M_sigma0"), &R.D("M_sigma1"), &R.D("M_sigma2") };

    for (int cpt=0; cpt < R.M_num_elements; ++cpt)
    {
        bool to_be_updated=true;
        if(std::binary_search(M_neumann_flags.begin(),M_neumann_flags.end(),idx[3*cpt+0]-1) ||
           std::binary_search(M_neumann_flags.begin(),M_neumann_flags.end(),idx[3*cpt+1]-1) ||
           std::binary_search(M_neumann_flags.begin(),M_neumann_flags.end(),idx[3*cpt+2]-1))
            to_be_updated=false;

        D_del_ci_ridge_myi[cpt] = 0.;

        double const surface_old = M_surface[cpt];
        double const old_conc = M_conc[cpt];
        M_surface[cpt] = measure(R, cpt);
        if((M_conc[cpt]>0.)  && (to_be_updated))
        {
            double const surf_ratio = surface_old/M_surface[cpt];
            M_conc[cpt] *= surf_ratio;
            M_thick[cpt] *= surf_ratio;
            M_snow_thick[cpt] *= surf_ratio;
            M_thick_myi[cpt]  *= surf_ratio;

            for(int k=0; k<3; k++)
                (*M_sigma[k])[cpt] *= surf_ratio;

            M_ridge_ratio[cpt] = 1. - (1.-M_ridge_ratio[cpt])*std::min(1., M_conc[cpt])/(old_conc*surf_ratio);

            if(young)
            {
                M_h_young[cpt] *= surf_ratio;
                M_conc_young[cpt] *= surf_ratio;
                M_hs_young[cpt] *= surf_ratio;
            }
            if ( equal_ridging )
            {
                double const conc_ratio = std::min(1.,M_conc[cpt])/old_conc;
                M_conc_myi[cpt] *= conc_ratio;
                D_del_ci_ridge_myi[cpt] = 0.;
            }
            else
            {
                M_conc_myi[cpt] *= surf_ratio;
                D_del_ci_ridge_myi[cpt] = -M_conc_myi[cpt];
                M_conc_myi[cpt] = std::min(M_conc_myi[cpt], 1.);
                D_del_ci_ridge_myi[cpt] += M_conc_myi[cpt];
            }
            D_del_ci_ridge_myi[cpt]*=days_in_sec/P.dtime_step;
        }

        double open_water_concentration=1.-M_conc[cpt];

        if ( young )
            open_water_concentration -= M_conc_young[cpt];

        open_water_concentration=(open_water_concentration<0.)?0.:open_water_concentration;
        open_water_concentration=(open_water_concentration>1.)?1.:open_water_concentration;

        double new_conc_young=0.;
        double new_h_young=0.;
        double new_hs_young=0.;

        double newice = 0.;
        double del_c = 0.;
        double newsnow = 0.;

        double ridge_young_ice_aspect_ratio=10.;

        if ( young )
        {
            if(M_conc_young[cpt]>0. )
            {
                new_conc_young   = std::min(1., std::max(0., 1. - M_conc[cpt] - open_water_concentration));

                if( (M_conc[cpt] > P.min_c) && (M_thick[cpt] > P.min_h) && (new_conc_young < M_conc_young[cpt] ))
                {
                    new_h_young      = new_conc_young*M_h_young[cpt]/M_conc_young[cpt];
                    new_hs_young     = new_conc_young*M_hs_young[cpt]/M_conc_young[cpt];

                    newice = M_h_young[cpt]-new_h_young;
                    del_c   = (M_conc_young[cpt]-new_conc_young)/ridge_young_ice_aspect_ratio;
                    newsnow = M_hs_young[cpt]-new_hs_young;

                    M_h_young[cpt]   = new_h_young;
                    M_hs_young[cpt]  = new_hs_young;

                    M_ridge_ratio[cpt] = 1. - (1.-M_ridge_ratio[cpt])*M_thick[cpt]/(M_thick[cpt]+newice);
                    M_thick[cpt]        += newice;
                    M_snow_thick[cpt]   += newsnow;
                }
            }
            else
            {
                M_h_young[cpt]=0.;
                M_hs_young[cpt]=0.;
            }
        }

        M_conc[cpt] = std::min(1.,std::max(0., 1. - new_conc_young - open_water_concentration + del_c));
        if ( young )
        {
            new_conc_young = std::max(0., std::min(new_conc_young, 1.- M_conc[cpt]));
            M_conc_young[cpt] = new_conc_young;
        }

        double max_true_thickness = 50.;
        if(M_conc[cpt]>0.)
        {
            double test_h_thick=M_thick[cpt]/M_conc[cpt];
            test_h_thick = (test_h_thick>max_true_thickness) ? max_true_thickness : test_h_thick ;
            M_conc[cpt]=std::min(1. - new_conc_young, M_thick[cpt]/test_h_thick);
        }
        else
        {
            M_ridge_ratio[cpt]=0.;
            M_thick[cpt]=0.;
            M_snow_thick[cpt]=0.;
        }

        M_conc[cpt]         = ((M_conc[cpt]>0.)?(M_conc[cpt] ):(0.)) ;
        M_thick[cpt]        = ((M_thick[cpt]>0.)?(M_thick[cpt]     ):(0.)) ;
        M_thick_myi[cpt]    = ((M_thick_myi[cpt]>0.)?(M_thick_myi[cpt]  ):(0.)) ;
        M_snow_thick[cpt]   = ((M_snow_thick[cpt]>0.)?(M_snow_thick[cpt]):(0.)) ;
        D_del_ci_ridge_myi[cpt] = -M_conc_myi[cpt];
        if (newice_type == 4 && use_young_ice_in_myi_reset == true)
            M_conc_myi[cpt] = std::max(0.,std::min(M_conc_myi[cpt],M_conc[cpt]+M_conc_young[cpt]));
        else
            M_conc_myi[cpt] = std::max(0.,std::min(M_conc_myi[cpt],M_conc[cpt]));
        D_del_ci_ridge_myi[cpt]+=M_conc_myi[cpt];
    }
}

// ---- bamg connectivity tables  contrib/bamg/src/Mesh.cpp:514-543, 583-629, 798-865 -----------
static void bamgTables(Rank& R)
{
    int const nbv = R.M_num_nodes;
    int const nbt = R.M_num_elements;
    auto const& idx = R.I("indices");      // BamgConvertMeshx input: M_mesh.indexTr(), 1-based

    // chaining algorithm, node -> (triangle,vertex) by head insertion  (526-537)
    std::vector<int> head_1(nbv,-1), next_1(3*(size_t)nbt), connectivitysize_1(nbv,0);
    int k=0;
    for (int i=0;i<nbt;i++)
        for (int j=0;j<3;j++)
        {
            int v=idx[3*i+j]-1;
            next_1[k]=head_1[v];
            head_1[v]=k++;
            connectivitysize_1[v]+=1;
        }
    int connectivitymax_1=0;
    for (int i=0;i<nbv;i++)
        if (connectivitysize_1[i]>connectivitymax_1) connectivitymax_1=connectivitysize_1[i];

    // NodalElementConnectivity (798-811), NaN padded doubles
    R.nec_width = connectivitymax_1;
    auto& NEC = R.D("NodalElementConnectivity");
    NEC.assign((size_t)connectivitymax_1*nbv, NAN);
    for (int i=0;i<nbv;i++)
    {
        k=0;
        for(int j=head_1[i];j!=-1;j=next_1[j])
        {
            NEC[(size_t)connectivitymax_1*i+k]=std::floor((double)j/3)+1;
            k++;
        }
    }

    // IssmEdges (583-629): edges numbered by first appearance, triangle order, local edges
    // VerticesOfTriangularEdge = {{1,2},{2,0},{0,1}} (contrib/bamg/include/macros.h:13)
    static const short VOTE[3][2] = {{1,2},{2,0},{0,1}};
    std::map<std::pair<int,int>,int> edge4;
    std::vector<int> e_i, e_j, e_first;   // sorted endpoints, first triangle
    for (int i=0;i<nbt;i++)
        for (int j=0;j<3;j++)
        {
            int i1=idx[3*i+VOTE[j][0]]-1;
            int i2=idx[3*i+VOTE[j][1]]-1;
            std::pair<int,int> key(std::min(i1,i2), std::max(i1,i2));
            if (edge4.find(key)==edge4.end())
            {
                edge4[key]=(int)e_i.size();
                e_i.push_back(key.first); e_j.push_back(key.second); e_first.push_back(i);
            }
        }
    int const ne = (int)e_i.size();
    std::vector<int> IssmEdges0(ne), IssmEdges1(ne);     // 1-based
    for (int i=0;i<ne;i++)
    {
        bool found=false;
        int const t = e_first[i];
        for (int j=0;j<3;j++)
        {
            if (idx[3*t+j]-1==e_i[i])
            {
                if (idx[3*t+(j+1)%3]-1==e_j[i]) { IssmEdges0[i]=e_i[i]+1; IssmEdges1[i]=e_j[i]+1; }
                else                           { IssmEdges0[i]=e_j[i]+1; IssmEdges1[i]=e_i[i]+1; }
                found=true;
                break;
            }
        }
        if (!found) throw std::logic_error("bamgTables: edge not found in its first triangle");
    }

    // NodalConnectivity (813-865): chain over edge endpoints, head insertion
    std::vector<int> head_2(nbv,-1), next_2(2*(size_t)ne), connectivitysize_2(nbv,0);
    k=0;
    for (int i=0;i<ne;i++)
        for (int j=0;j<2;j++)
        {
            int v=(j==0?IssmEdges0[i]:IssmEdges1[i])-1;
            next_2[k]=head_2[v];
            head_2[v]=k++;
            connectivitysize_2[v]+=1;
        }
    int connectivitymax_2=0;
    for (int i=0;i<nbv;i++)
        if (connectivitysize_2[i]>connectivitymax_2) connectivitymax_2=connectivitysize_2[i];
    connectivitymax_2++;
    R.nc_width = connectivitymax_2;
    auto& NC = R.D("NodalConnectivity");
    NC.assign((size_t)connectivitymax_2*nbv, 0.);
    for (int i=0;i<nbv;i++)
    {
        k=0;
        for(int j=head_2[i];j!=-1;j=next_2[j])
        {
            int num=IssmEdges0[j/2];
            if (i+1==num)
                NC[(size_t)connectivitymax_2*i+k]=IssmEdges1[j/2];
            else
                NC[(size_t)connectivitymax_2*i+k]=num;
            k++;
        }
        NC[(size_t)connectivitymax_2*(i+1)-1]=k;
    }
}

// ---- nodalGrid  core/src/gmshmesh.cpp:856-1498 + element load filter entities.hpp:105-134 ----
struct GElt { int number; int partition; std::vector<int> ghosts; bool is_ghost; int indices[3]; int ghostNodes[3]; };

static void nodalGridAll(int const nranks, int const global_num_nodes, double const* gx, double const* gy,
                         int const global_num_elements, int const* gtri /*1-based*/, int const* gpart,
                         int const* ghost_ptr, int const* ghost_val, std::vector<Rank*>& W)
{
    // per-rank state that the reference keeps as GmshMesh members
    struct M {
        std::vector<GElt> M_triangles;
        std::vector<int> M_local_dof_without_ghost, M_local_ghost, M_local_dof_with_ghost;
        std::vector<int> M_triangles_id_with_ghost;
        std::map<int,int> M_transfer_map, M_transfer_map_reordered;
        std::map<int,int> reorder;
        std::vector<int> triangles_num_without_ghost;
        int M_num_triangles_without_ghost = 0;
    };
    std::vector<M> S(nranks);

    // readFromFile: keep the elements that are "on processor" (entities.hpp:105-134)
    for (int r=0; r<nranks; ++r)
    {
        for (int e=0; e<global_num_elements; ++e)
        {
            // setPartition() (entities.hpp:105-134) decides from the tags alone; build the element only if it stays
            bool on_proc = false, is_ghost = false;
            if (nranks == 1) on_proc = true;
            else if (r == gpart[e]) on_proc = true;
            else
                for (int q=ghost_ptr[e]; q<ghost_ptr[e+1]; ++q)
                    if (ghost_val[q] % nranks == r) { on_proc = true; is_ghost = true; break; }
            if (!on_proc) continue;
            GElt g;
            g.number = e+1;
            g.partition = gpart[e];
            for (int q=ghost_ptr[e]; q<ghost_ptr[e+1]; ++q)
                g.ghosts.push_back(ghost_val[q] % nranks);
            for (int i=0;i<3;++i) { g.indices[i] = gtri[3*e+i]; g.ghostNodes[i]=0; }
            g.is_ghost = is_ghost;
            S[r].M_triangles.push_back(g);
        }
    }

    // 862-979 on every rank
    for (int r=0; r<nranks; ++r)
    {
        M& m = S[r];
        std::vector<int> ghosts_nodes_f;
        for (auto it=m.M_triangles.begin(); it!=m.M_triangles.end(); ++it)
            if (it->is_ghost)
                for (int const& index : it->indices)
                    ghosts_nodes_f.push_back(index);
        std::sort(ghosts_nodes_f.begin(), ghosts_nodes_f.end());
        ghosts_nodes_f.erase(std::unique( ghosts_nodes_f.begin(), ghosts_nodes_f.end() ), ghosts_nodes_f.end());

        for (auto it=m.M_triangles.begin(); it!=m.M_triangles.end(); ++it)
        {
            if (!it->is_ghost)
            {
                for (int const& index : it->indices)
                {
                    if (!std::binary_search(ghosts_nodes_f.begin(),ghosts_nodes_f.end(),index))
                        m.M_local_dof_without_ghost.push_back(index);

                    if ((it->ghosts.size() > 0) && (std::binary_search(ghosts_nodes_f.begin(),ghosts_nodes_f.end(),index)))
                        m.M_local_dof_without_ghost.push_back(index);
                }
            }
        }
        std::sort(m.M_local_dof_without_ghost.begin(), m.M_local_dof_without_ghost.end());
        m.M_local_dof_without_ghost.erase(std::unique( m.M_local_dof_without_ghost.begin(), m.M_local_dof_without_ghost.end() ), m.M_local_dof_without_ghost.end());

        std::vector<int> all_local_nodes;
        for (auto it=m.M_triangles.begin(); it!=m.M_triangles.end(); ++it)
        {
            if (r <= it->partition)
            {
                bool is_found = false;
                for (int i=0; i<3; ++i)
                    if (std::binary_search(m.M_local_dof_without_ghost.begin(), m.M_local_dof_without_ghost.end(),it->indices[i]))
                    {
                        is_found = true;
                        break;
                    }
                if (!is_found)
                    continue;
                for (int i=0; i<3; ++i)
                    all_local_nodes.push_back(it->indices[i]);
            }
        }
        std::sort(all_local_nodes.begin(), all_local_nodes.end());
        all_local_nodes.erase(std::unique( all_local_nodes.begin(), all_local_nodes.end() ), all_local_nodes.end());
        std::set_difference(all_local_nodes.begin(), all_local_nodes.end(),
                            m.M_local_dof_without_ghost.begin(), m.M_local_dof_without_ghost.end(),
                            std::back_inserter(m.M_local_ghost));
    }

    // allGather (1047): renumbering[ii] = rank ii's provisional owned list
    std::vector<std::vector<int>> renumbering(nranks);
    int num_nodes = 0;
    for (int r=0; r<nranks; ++r) { renumbering[r] = S[r].M_local_dof_without_ghost; num_nodes += (int)renumbering[r].size(); }

    // 1057-1094: every rank runs the same de-duplication; the lowest rank keeps shared nodes
    if (global_num_nodes != num_nodes)
    {
        if (global_num_nodes < num_nodes)
        {
            for (int ii=0; ii<nranks; ++ii)
                for (int jj=0; jj<nranks; ++jj)
                    if (ii != jj)
                    {
                        std::vector<int> duplicated_dofs;
                        std::set_intersection(renumbering[ii].begin(),renumbering[ii].end(),
                                              renumbering[jj].begin(),renumbering[jj].end(),
                                              std::back_inserter(duplicated_dofs));
                        if (duplicated_dofs.size() == 0)
                            continue;
                        if (jj < ii)
                        {
                            // erase(remove(x)) for every duplicated x, done as one sorted difference
                            std::vector<int> kept;
                            std::set_difference(renumbering[ii].begin(),renumbering[ii].end(),
                                                duplicated_dofs.begin(),duplicated_dofs.end(),
                                                std::back_inserter(kept));
                            renumbering[ii].swap(kept);
                            for (int dd : duplicated_dofs)
                                S[ii].M_local_ghost.push_back(dd);
                        }
                    }
            for (int r=0; r<nranks; ++r)
                S[r].M_local_dof_without_ghost = renumbering[r];
        }
    }

    // 1165-1220: global renumbering, contiguous per rank, u block then v block
    std::map<int,int> reorder;
    {
        int cpts = 0, cpts_dom = 0;
        for (int ii=0; ii<nranks; ++ii)
        {
            int sr = (int)renumbering[ii].size();
            for (int jj=0; jj<sr; ++jj)
            {
                reorder.insert(std::make_pair(renumbering[ii][jj],cpts+1+cpts_dom));
                reorder.insert(std::make_pair(renumbering[ii][jj]+global_num_nodes,cpts+1+sr+cpts_dom));
                ++cpts;
            }
            cpts_dom += (int)renumbering[ii].size();
        }
    }

    std::vector<std::vector<int>> tri_renumbering(nranks);
    for (int r=0; r<nranks; ++r)
    {
        M& m = S[r];
        Rank& R = *W[r];
        R.rank = r; R.nranks = nranks;
        std::sort(m.M_local_dof_without_ghost.begin(), m.M_local_dof_without_ghost.end());
        std::sort(m.M_local_ghost.begin(), m.M_local_ghost.end());
        m.M_local_dof_with_ghost = m.M_local_dof_without_ghost;
        m.M_local_dof_with_ghost.insert(m.M_local_dof_with_ghost.end(), m.M_local_ghost.begin(), m.M_local_ghost.end());

        auto local_dof_with_ghost = m.M_local_dof_with_ghost;       // M_local_dof_with_ghost_init
        int const M_nldof_with_ghost = (int)m.M_local_dof_with_ghost.size();
        int const M_nldof_without_ghost = (int)m.M_local_dof_without_ghost.size();

        auto& coordX = R.D("coordX"); coordX.resize(M_nldof_with_ghost);
        auto& coordY = R.D("coordY"); coordY.resize(M_nldof_with_ghost);
        std::vector<int> dof_with_ghost_reordered(2*(size_t)M_nldof_with_ghost);
        for (int k=0; k<(int)local_dof_with_ghost.size(); ++k)
        {
            m.M_transfer_map.insert(std::make_pair(local_dof_with_ghost[k],k+1));
            int rdof = reorder.find(local_dof_with_ghost[k])->second;
            int rdofv = reorder.find(local_dof_with_ghost[k]+global_num_nodes)->second;
            m.M_transfer_map_reordered.insert(std::make_pair(rdof,k+1));
            dof_with_ghost_reordered[k] = rdof;
            dof_with_ghost_reordered[k+M_nldof_with_ghost] = rdofv;
            coordX[k] = gx[local_dof_with_ghost[k]-1];
            coordY[k] = gy[local_dof_with_ghost[k]-1];
            if (k >= M_nldof_without_ghost)
                m.M_local_ghost[k-M_nldof_without_ghost] = rdof;
        }
        std::sort(m.M_local_ghost.begin(), m.M_local_ghost.end());

        R.M_num_nodes = M_nldof_with_ghost;
        R.M_local_ndof = M_nldof_without_ghost;
        R.I("local_dof_with_ghost_init") = local_dof_with_ghost;       // global (file) ids, 1-based
        R.I("local_dof_with_ghost") = dof_with_ghost_reordered;        // reordered global ids
        R.I("local_ghost") = m.M_local_ghost;                          // reordered ids, sorted

        // 1267-1312
        std::vector<GElt> _triangles = m.M_triangles;
        m.M_triangles.resize(0);
        for (auto it=_triangles.begin(); it!=_triangles.end(); ++it)
        {
            if (r <= it->partition)
            {
                bool _test = false;
                for (int i=0; i<3; ++i)
                    if (m.M_transfer_map.find(it->indices[i]) == m.M_transfer_map.end())
                    {
                        _test = true;
                        break;
                    }
                if (_test)
                    continue;
                for (int i=0; i<3; ++i)
                {
                    int rdof = reorder.find(it->indices[i])->second;
                    it->indices[i] = m.M_transfer_map.find(it->indices[i])->second;
                    it->ghostNodes[i] = std::binary_search(m.M_local_ghost.begin(),m.M_local_ghost.end(),rdof) ? 1 : 0;
                }
                m.M_triangles.push_back(*it);
                if (r == it->partition)
                    m.triangles_num_without_ghost.push_back(it->number);
            }
        }
        tri_renumbering[r] = m.triangles_num_without_ghost;
    }

    // 1316-1373: detect elements that no rank owns
    int num_elements = 0;
    for (int r=0; r<nranks; ++r) num_elements += (int)tri_renumbering[r].size();
    std::vector<int> diff_trs;
    if (global_num_elements != num_elements)
    {
        std::vector<int> all_trs;
        for (int ii=0; ii<nranks; ++ii)
            for (int v : tri_renumbering[ii]) all_trs.push_back(v);
        std::sort(all_trs.begin(), all_trs.end());
        std::vector<int> global_trs(all_trs.size());
        std::iota(global_trs.begin(), global_trs.end(), 1);
        std::set_difference(global_trs.begin(), global_trs.end(), all_trs.begin(), all_trs.end(),
                            std::back_inserter(diff_trs));
    }

    // 1377-1423: owned elements first, ghosts after
    for (int r=0; r<nranks; ++r)
    {
        M& m = S[r];
        Rank& R = *W[r];
        std::vector<GElt> _triangles = m.M_triangles;
        m.M_triangles.resize(0);
        for (auto it=_triangles.begin(); it!=_triangles.end(); ++it)
        {
            for (int i=0; i<(int)diff_trs.size(); ++i)
                if (it->number == diff_trs[i])
                {
                    int min_rank = *std::min_element(it->ghosts.begin(), it->ghosts.end());
                    if (r == min_rank)
                    {
                        m.triangles_num_without_ghost.push_back(it->number);
                        it->partition = r;
                    }
                }
            if (r == it->partition)
            {
                m.M_triangles.push_back(*it);
                m.M_triangles_id_with_ghost.push_back(it->number);
                ++m.M_num_triangles_without_ghost;
            }
        }
        for (auto it=_triangles.begin(); it!=_triangles.end(); ++it)
            if (r != it->partition)
            {
                m.M_triangles.push_back(*it);
                m.M_triangles_id_with_ghost.push_back(it->number);
            }

        R.M_num_elements = (int)m.M_triangles.size();
        R.M_local_nelements = m.M_num_triangles_without_ghost;
        auto& indices = R.I("indices");       indices.resize(3*(size_t)R.M_num_elements);
        auto& ghostNodes = R.I("ghostNodes"); ghostNodes.resize(3*(size_t)R.M_num_elements);
        auto& epart = R.I("element_partition"); epart.resize(R.M_num_elements);
        for (int e=0; e<R.M_num_elements; ++e)
        {
            for (int i=0;i<3;++i) { indices[3*e+i] = m.M_triangles[e].indices[i]; ghostNodes[3*e+i] = m.M_triangles[e].ghostNodes[i]; }
            epart[e] = m.M_triangles[e].partition;
        }
        R.I("triangles_id_with_ghost") = m.M_triangles_id_with_ghost;   // file element numbers, 1-based
    }

    // initUpdateGhosts FE.cpp:14003-14088 (M_sizes_nodes = owned node counts, gatherSizes 2015-2036)
    std::vector<int> M_sizes_nodes(nranks);
    for (int r=0; r<nranks; ++r) M_sizes_nodes[r] = W[r]->M_local_ndof;
    auto globalNumToprocId = [&](int global_num) {            // 14093-14105
        int cpt = 0;
        for (int i=0; i<nranks; i++)
        {
            if ((cpt < global_num) && (global_num <= cpt+M_sizes_nodes[i]))
                return i;
            cpt += 2*M_sizes_nodes[i];
        }
        throw std::logic_error("Couldn't map global number to proc id");
    };
    std::vector<std::vector<std::vector<int>>> local_ghosts_global_index(nranks, std::vector<std::vector<int>>(nranks));
    for (int r=0; r<nranks; ++r)
    {
        M& m = S[r];
        Rank& R = *W[r];
        for (int currentid : m.M_local_ghost)
            local_ghosts_global_index[r][globalNumToprocId(currentid)].push_back(currentid);
        R.M_local_ghosts_proc_id.resize(0);
        R.M_local_ghosts_local_index.assign(nranks, std::vector<int>());
        for (int i=0; i<nranks; i++)
        {
            if (local_ghosts_global_index[r][i].size() != 0)
                R.M_local_ghosts_proc_id.push_back(i);
            R.M_local_ghosts_local_index[i].resize(local_ghosts_global_index[r][i].size());
            for (int j=0; j<(int)local_ghosts_global_index[r][i].size(); j++)
                R.M_local_ghosts_local_index[i][j] = m.M_transfer_map_reordered.find(local_ghosts_global_index[r][i][j])->second-1;
        }
    }
    for (int r=0; r<nranks; ++r)
    {
        M& m = S[r];
        Rank& R = *W[r];
        R.M_recipients_proc_id.resize(0);
        for (int i=0; i<nranks; i++)
            for (int p : W[i]->M_local_ghosts_proc_id)
                if (p == r)
                    R.M_recipients_proc_id.push_back(i);
        R.M_extract_local_index.assign(nranks, std::vector<int>());
        for (int proc : R.M_recipients_proc_id)
        {
            auto const& extract_global_index = local_ghosts_global_index[proc][r];
            R.M_extract_local_index[proc].resize(extract_global_index.size());
            for (int j=0; j<(int)extract_global_index.size(); j++)
                R.M_extract_local_index[proc][j] = m.M_transfer_map_reordered.find(extract_global_index[j])->second-1;
        }
    }
}

// single-rank mesh (the reference skips nodalGrid when comm.size()==1, gmshmesh.cpp:133-134;
// the natural 1-partition limit is "every node and element owned, local id == file id")
static void singleRankMesh(Rank& R, int nn, double const* gx, double const* gy, int ne, int const* gtri)
{
    R.rank = 0; R.nranks = 1;
    R.M_num_nodes = nn; R.M_local_ndof = nn; R.M_num_elements = ne; R.M_local_nelements = ne;
    R.D("coordX").assign(gx, gx+nn);
    R.D("coordY").assign(gy, gy+nn);
    R.I("indices").assign(gtri, gtri+3*(size_t)ne);
    R.I("ghostNodes").assign(3*(size_t)ne, 0);
    auto& a = R.I("local_dof_with_ghost_init"); a.resize(nn); std::iota(a.begin(), a.end(), 1);
    auto& b = R.I("triangles_id_with_ghost"); b.resize(ne); std::iota(b.begin(), b.end(), 1);
    R.I("element_partition").assign(ne, 0);
    R.M_extract_local_index.assign(1, {}); R.M_local_ghosts_local_index.assign(1, {});
}

// bcMarkedNodes FE.cpp:150-271 : flags_root holds 1-based file node ids (FE.cpp:323-333)
static void bcMarkedNodes(Rank& R, int const* flags_root, int dir_size, int nmn_size)
{
    auto const& init = R.I("local_dof_with_ghost_init");
    std::map<int,int> transfer;
    for (int k=0;k<(int)init.size();++k) transfer.insert(std::make_pair(init[k],k+1));
    std::vector<int> M_dirichlet_flags, M_neumann_flags;
    auto& M_mask = R.I("M_mask");                     M_mask.assign(R.M_num_nodes,0);
    auto& M_mask_dirichlet = R.I("M_mask_dirichlet"); M_mask_dirichlet.assign(R.M_num_nodes,0);
    for (int i=0; i<dir_size+nmn_size; ++i)
    {
        auto f = transfer.find(flags_root[i]);
        if (f != transfer.end())
        {
            int lindex = f->second-1;
            if (i < dir_size)
            {
                if (lindex < R.M_local_ndof)
                {
                    M_dirichlet_flags.push_back(lindex);
                    M_mask_dirichlet[lindex] = 1;
                }
            }
            else
                M_neumann_flags.push_back(lindex);
            M_mask[lindex] = 1;
        }
    }
    std::sort(M_dirichlet_flags.begin(), M_dirichlet_flags.end());
    M_dirichlet_flags.erase(std::unique(M_dirichlet_flags.begin(), M_dirichlet_flags.end() ), M_dirichlet_flags.end());
    std::sort(M_neumann_flags.begin(), M_neumann_flags.end());
    M_neumann_flags.erase(std::unique(M_neumann_flags.begin(), M_neumann_flags.end() ), M_neumann_flags.end());
    auto& dn = R.I("M_dirichlet_nodes"); dn.resize(2*M_dirichlet_flags.size());
    for (int i=0; i<(int)M_dirichlet_flags.size(); ++i)
    {
        dn[2*i] = M_dirichlet_flags[i];
        dn[2*i+1] = M_dirichlet_flags[i]+R.M_num_nodes;
    }
    auto& nn = R.I("M_neumann_nodes"); nn.resize(2*M_neumann_flags.size());
    for (int i=0; i<(int)M_neumann_flags.size(); ++i)
    {
        nn[2*i] = M_neumann_flags[i];
        nn[2*i+1] = M_neumann_flags[i]+R.M_num_nodes;
    }
    R.I("M_dirichlet_flags") = M_dirichlet_flags;
    R.I("M_neumann_flags") = M_neumann_flags;
}

// ---- section 8(f) row 1: checkRegridding / updateIceDiagnostics --------------------------------------
// FE.cpp:1758-1768  minAngles(element, mesh, um, factor=1)
static double minAngles(Rank& R, int cpt)
{
    double side[3];
    sides(R, cpt, side);
    std::sort(side, side+3);
    double minang = std::acos( (std::pow(side[1],2.) + std::pow(side[2],2.) - std::pow(side[0],2.) )/(2*side[1]*side[2]) );
    minang = minang*45.0/std::atan(1.0);
    return minang;
}

// FE.cpp:8298-8309 without the final all_reduce (one bool; the caller ORs the ranks)
//   out[0] = minAngle(M_mesh, M_UM, 1.) of THIS rank (FE.cpp:1795-1806, before the all_reduce at 1812)
//   out[1], out[2] = min / max jacobian of flip() (FE.cpp:1824-1839)
//   flags[0] = flip(), flags[1] = regrid_local
static void checkRegriddingLocal(Rank& R, double const regrid_angle, double* out, int* flags)
{
    int const ne = R.M_num_elements;       // M_mesh.triangles(): owned + ghost elements
    std::vector<double> all_min_angle(ne), area(ne);
    for (int cpt=0; cpt<ne; ++cpt)
    {
        all_min_angle[cpt] = minAngles(R, cpt);
        double v[3][2];
        vertices(R, cpt, v);
        area[cpt] = jacobian(v);
    }
    double const minang = *std::min_element(all_min_angle.begin(), all_min_angle.end());
    double const minarea = *std::min_element(area.begin(), area.end());
    double const maxarea = *std::max_element(area.begin(), area.end());
    bool const flipped = ((minarea <= 0.) && (maxarea >= 0.));
    out[0] = minang; out[1] = minarea; out[2] = maxarea;
    flags[0] = flipped;
    flags[1] = (minang < regrid_angle) || flipped;
}

// FE.cpp:7860-7900.  D_tsurf needs M_tice / M_sst / M_tsurf_young (thermodynamics state that never
// enters the path) and is left to the host; D_dmean / D_dmax are constants.
static void updateIceDiagnostics(Rank& R, OrcParams const& P)
{
    int const ne = R.M_num_elements, nn = R.M_num_nodes;
    auto& D_conc = R.D("D_conc"); D_conc.resize(ne);
    auto& D_thick = R.D("D_thick"); D_thick.resize(ne);
    auto& D_snow_thick = R.D("D_snow_thick"); D_snow_thick.resize(ne);
    auto& D_sigma0 = R.D("D_sigma0"); D_sigma0.resize(ne);
    auto& D_sigma1 = R.D("D_sigma1"); D_sigma1.resize(ne);
    auto& D_divergence = R.D("D_divergence"); D_divergence.resize(ne);
    auto const& M_conc = R.D("M_conc"); auto const& M_thick = R.D("M_thick"); auto const& M_snow_thick = R.D("M_snow_thick");
    auto const& M_conc_young = R.D("M_conc_young"); auto const& M_h_young = R.D("M_h_young"); auto const& M_hs_young = R.D("M_hs_young");
    auto const& s0 = R.D("M_sigma0"); auto const& s1 = R.D("M_sigma1"); auto const& s2 = R.D("M_sigma2");
    auto const& M_VT = R.D("M_VT");
    auto const& idx = R.I("indices");
    for (int i=0; i<ne; ++i)
    {
        D_conc[i] = M_conc[i];
        D_thick[i] = M_thick[i];
        D_snow_thick[i] = M_snow_thick[i];
        if (P.ice_cat_type == 1)
        {
            D_conc[i] += M_conc_young[i];
            D_thick[i] += M_h_young[i];
            D_snow_thick[i] += M_hs_young[i];
        }
        D_sigma0[i] =            (s0[i]+s1[i])/2.;
        D_sigma1[i] = std::hypot((s0[i]-s1[i])/2., s2[i]);

        D_divergence[i] = 0.;
        double shape_coeff[6];
        shapeCoeff(R, i, shape_coeff);
        for (int j=0; j<3; ++j)
        {
            double const u = M_VT[idx[3*i+j]-1];
            double const v = M_VT[idx[3*i+j]-1 + nn];
            double const dxN = shape_coeff[j];
            double const dyN = shape_coeff[j+3];
            D_divergence[i] += dxN*u + dyN*v;
        }
    }
}

// ---- section 8(f) row 2: ExternalData::getVector()  externaldata.cpp:366-455 ----------------------------
// d0 / d1 = Dataset::variables[...].interpolated_data[0/1] of the variable (scalars) or of its two
// components concatenated [u | v] (vectors: get(i) picks component i>=M_target_size, :386-395).
static void externalDataGetVector(long n, double const* d0, double const* d1, int interp_linear_time,
                                  double M_current_time, double const ftime_range[2], double M_factor,
                                  double M_bias_correction, double* out)
{
    double fcoeff[2] = {1., 0.};
    if (interp_linear_time)
    {
        double const fdt = std::abs(ftime_range[1]-ftime_range[0]);
        fcoeff[0] = std::abs(M_current_time-ftime_range[1])/fdt;
        fcoeff[1] = std::abs(M_current_time-ftime_range[0])/fdt;
    }
    for (long i=0; i<n; ++i)
    {
        double value;
        if (interp_linear_time)
            value = M_factor*(fcoeff[0]*d0[i] + fcoeff[1]*d1[i]);
        else
            value = M_factor*d0[i];
        out[i] = value + M_bias_correction;
    }
}

} // namespace

// =====================================================================================
// C interface (ctypes)
// =====================================================================================
static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (std::exception const& e) { g_err = e.what(); return -1; } return 0;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// Defaults of the options the path reads, restated from /root/reference/model/options.cpp (line numbers in the
// comments) so that the CPU arm of bench.py needs nothing from the product library.  cohesion[0..2] receives
// dynamics.C_lab, dynamics.alea_factor, dynamics.time_relaxation_damage (host-side recipe, FE.cpp:6995-6999).
void orc_params_defaults(OrcParams* p, double* cohesion)
{
    std::memset(p, 0, sizeof(*p));
    p->dynamics_type = 0;                                  // setup.dynamics-type = bbm            options.cpp:111
    p->basal_stress_type = 1;                              // setup.basal_stress-type = lemieux    options.cpp:109
    p->newice_type = 4;                                    // thermo.newice_type                   options.cpp:397
    p->ice_cat_type = 1;                                   // newice_type == 4 -> YOUNG_ICE        FE.cpp:1212-1215
    p->substeps = 120;                                     // dynamics.substeps                    options.cpp:363
    p->equal_ridging = 0;                                  // age.equal_ridging                    options.cpp:547
    p->use_young_ice_in_myi_reset = 1;                     // age.include_young_ice                options.cpp:545
    p->dtime_step = 200.;                                  // simul.timestep                       options.cpp:43
    p->ocean_turning_angle_rad = (3.14159265358979323846/180.)*25.;   // dynamics.oceanic_turning_angle   options.cpp:347, FE.cpp:1172
    p->min_h = 0.05; p->min_c = 0.01;                      // options.cpp:325-326
    p->young = 5.9605e+08;                                 // options.cpp:313
    p->nu0 = 1./3.;                                        // options.cpp:318
    p->tan_phi = 0.7;                                      // options.cpp:319
    p->compr_strength = 1e10;                              // options.cpp:320 (the host multiplies by scale_coef, FE.cpp:6998)
    p->compaction_param = -20.;                            // options.cpp:321
    p->undamaged_time_relaxation_sigma = 1e7;              // options.cpp:331
    p->exponent_relaxation_sigma = 5.;                     // options.cpp:333
    p->compression_factor = 10e3;                          // options.cpp:359
    p->exponent_compression_factor = 1.5;                  // options.cpp:358
    p->quad_drag_coef_water = 0.0055;                      // options.cpp:342
    p->evp_e = 2.; p->evp_Pstar = 27.5e3; p->evp_C = 20.; p->evp_dmin = 1e-9;   // options.cpp:365-372
    p->mevp_alpha = 500.; p->mevp_beta = 500.;             // options.cpp:375-376
    p->basal_k1 = 10.; p->basal_k2 = 15.; p->basal_Cb = 20.; p->basal_u0 = 5e-5;   // options.cpp:350-353
    if (cohesion) { cohesion[0] = 2.0e6; cohesion[1] = 0.; cohesion[2] = 25.; }     // options.cpp:317, 311, 329
}

void* orc_rank_create() { return new Rank(); }
void orc_rank_destroy(void* h) { delete (Rank*)h; }

int orc_rank_set_double(void* h, const char* name, const double* p, long n)
{ ORC_TRY ((Rank*)h)->d[name].assign(p, p+n); ORC_CATCH }
int orc_rank_set_int(void* h, const char* name, const int* p, long n)
{ ORC_TRY ((Rank*)h)->i[name].assign(p, p+n); ORC_CATCH }
long orc_rank_size_double(void* h, const char* name)
{ auto& m = ((Rank*)h)->d; auto f = m.find(name); return f==m.end() ? -1 : (long)f->second.size(); }
long orc_rank_size_int(void* h, const char* name)
{ auto& m = ((Rank*)h)->i; auto f = m.find(name); return f==m.end() ? -1 : (long)f->second.size(); }
int orc_rank_get_double(void* h, const char* name, double* p, long n)
{ ORC_TRY auto& v = ((Rank*)h)->d.at(name); if ((long)v.size()!=n) throw std::length_error(std::string("size mismatch for ")+name); std::copy(v.begin(), v.end(), p); ORC_CATCH }
int orc_rank_get_int(void* h, const char* name, int* p, long n)
{ ORC_TRY auto& v = ((Rank*)h)->i.at(name); if ((long)v.size()!=n) throw std::length_error(std::string("size mismatch for ")+name); std::copy(v.begin(), v.end(), p); ORC_CATCH }

// sizes: out[0..5] = M_num_nodes, M_local_ndof, M_num_elements, M_local_nelements, nec_width, nc_width
void orc_rank_sizes(void* h, int* out)
{
    Rank& R = *(Rank*)h;
    out[0]=R.M_num_nodes; out[1]=R.M_local_ndof; out[2]=R.M_num_elements; out[3]=R.M_local_nelements;
    out[4]=R.nec_width; out[5]=R.nc_width;
}

// halo lists: which = 0 -> M_extract_local_index[proc], 1 -> M_local_ghosts_local_index[proc]
long orc_rank_halo_size(void* h, int which, int proc)
{
    Rank& R = *(Rank*)h;
    auto& L = which==0 ? R.M_extract_local_index : R.M_local_ghosts_local_index;
    if (proc<0 || proc>=(int)L.size()) return 0;
    return (long)L[proc].size();
}
int orc_rank_halo_get(void* h, int which, int proc, int* out)
{
    ORC_TRY
    Rank& R = *(Rank*)h;
    auto& L = which==0 ? R.M_extract_local_index : R.M_local_ghosts_local_index;
    std::copy(L.at(proc).begin(), L.at(proc).end(), out);
    ORC_CATCH
}
int orc_rank_halo_set(void* h, int nranks, int which, int proc, const int* p, long n)
{
    ORC_TRY
    Rank& R = *(Rank*)h;
    auto& L = which==0 ? R.M_extract_local_index : R.M_local_ghosts_local_index;
    if ((int)L.size()!=nranks) L.assign(nranks, std::vector<int>());
    L.at(proc).assign(p, p+n);
    ORC_CATCH
}

int orc_single_rank_mesh(void* h, int nn, const double* gx, const double* gy, int ne, const int* gtri)
{ ORC_TRY singleRankMesh(*(Rank*)h, nn, gx, gy, ne, gtri); ORC_CATCH }

int orc_nodal_grid(int nranks, void** handles, int nn, const double* gx, const double* gy, int ne,
                   const int* gtri, const int* gpart, const int* ghost_ptr, const int* ghost_val)
{
    ORC_TRY
    std::vector<Rank*> W(nranks);
    for (int r=0;r<nranks;++r) W[r]=(Rank*)handles[r];
    nodalGridAll(nranks, nn, gx, gy, ne, gtri, gpart, ghost_ptr, ghost_val, W);
    ORC_CATCH
}

int orc_bamg_tables(void* h) { ORC_TRY bamgTables(*(Rank*)h); ORC_CATCH }

int orc_bc_marked_nodes(void* h, const int* flags_root, int dir_size, int nmn_size)
{ ORC_TRY bcMarkedNodes(*(Rank*)h, flags_root, dir_size, nmn_size); ORC_CATCH }

// FE.cpp:11459-11475 + 3909-3914: boost::minstd_rand (x <- 48271 x mod 2^31-1, seed 1) through
// boost::uniform_01 ((x-1)/2147483646), drawn in global element order, indexed by file element number
int orc_calc_cohesion(void* h, int global_num_elements, double C_fix, double C_alea)
{
    ORC_TRY
    Rank& R = *(Rank*)h;
    std::vector<double> random_number_root(global_num_elements);
    uint64_t x = 1;
    for (int i=0; i<global_num_elements; ++i)
    {
        x = (x*48271ULL) % 2147483647ULL;
        random_number_root[i] = double(x - 1) * (1.0/2147483646.0);
    }
    auto const& id_elements = R.I("triangles_id_with_ghost");
    auto& M_Cohesion = R.D("M_Cohesion"); M_Cohesion.resize(R.M_num_elements);
    auto& M_random_number = R.D("M_random_number"); M_random_number.resize(R.M_num_elements);
    for (int i=0; i<R.M_num_elements; ++i)
    {
        M_random_number[i] = random_number_root[id_elements[i]-1];
        M_Cohesion[i] = C_fix+C_alea*(M_random_number[i]);
    }
    ORC_CATCH
}

int orc_explicit_solve(int nranks, void** handles, const OrcParams* P)
{
    ORC_TRY
    std::vector<Rank*> W(nranks);
    for (int r=0;r<nranks;++r) W[r]=(Rank*)handles[r];
    explicitSolve(W, *P);
    ORC_CATCH
}

int orc_update(void* h, const OrcParams* P) { ORC_TRY update(*(Rank*)h, *P); ORC_CATCH }

int orc_check_regridding(void* h, double regrid_angle, double* out3, int* flags2)
{ ORC_TRY checkRegriddingLocal(*(Rank*)h, regrid_angle, out3, flags2); ORC_CATCH }

int orc_update_ice_diagnostics(void* h, const OrcParams* P) { ORC_TRY updateIceDiagnostics(*(Rank*)h, *P); ORC_CATCH }

int orc_external_data_get_vector(long n, const double* d0, const double* d1, int interp_linear_time, double current_time,
                                 double ftime0, double ftime1, double factor, double bias_correction, double* out)
{
    ORC_TRY
    double const ftime_range[2] = {ftime0, ftime1};
    externalDataGetVector(n, d0, d1, interp_linear_time, current_time, ftime_range, factor, bias_correction, out);
    ORC_CATCH
}

int orc_update_ghosts(int nranks, void** handles, const char* name)
{
    ORC_TRY
    std::vector<Rank*> W(nranks);
    for (int r=0;r<nranks;++r) W[r]=(Rank*)handles[r];
    updateGhosts(W, name);
    ORC_CATCH
}

// Timed variant for bench.py's cpu_baseline: runs prep once, then `nsub` sub-cycles, and returns the
// wall time (seconds) of the sub-cycle loop only (the reference's "sub-time stepping" timer).
double orc_time_subcycles(int nranks, void** handles, const OrcParams* P, int nsub);

} // extern "C"

#include <chrono>
extern "C" double orc_time_subcycles(int nranks, void** handles, const OrcParams* P, int nsub)
{
    try {
        std::vector<Rank*> W(nranks);
        for (int r=0;r<nranks;++r) W[r]=(Rank*)handles[r];
        double const dte = P->dtime_step/double(P->substeps);
        for (Rank* R : W) { initFETensors(*R, *P); explicitSolve_prep(*R, *P); }
        auto t0 = std::chrono::steady_clock::now();
        for (int s=0; s<nsub; ++s)
        {
            for (Rank* R : W) explicitSolve_substep(*R, *P, dte);
            updateGhosts(W, "M_VT");
            if ( P->dynamics_type != 4 )
                for (Rank* R : W) moveMesh(*R, dte);
        }
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1-t0).count();
    } catch (std::exception const& e) { g_err = e.what(); return -1.; }
}

// Multi-threaded timing variant: one std::thread per rank (the shape of the reference's
// "one MPI rank per core" run), ghost exchange through shared memory with barriers in place
// of the blocking send/recv pairs of FE.cpp:13981-13985.
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
namespace {
struct Barrier {
    std::mutex m; std::condition_variable cv; int n, count = 0, gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        int g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&]{ return g != gen; });
    }
};
}
extern "C" double orc_time_subcycles_mt(int nranks, void** handles, const OrcParams* P, int nsub)
{
    try {
        std::vector<Rank*> W(nranks);
        for (int r=0;r<nranks;++r) W[r]=(Rank*)handles[r];
        double const dte = P->dtime_step/double(P->substeps);
        std::vector<std::vector<std::vector<double>>> msg(nranks, std::vector<std::vector<double>>(nranks));
        Barrier bar(nranks);
        std::vector<double> secs(nranks, 0.);
        auto body = [&](int r) {
            Rank& R = *W[r];
            initFETensors(R, *P); explicitSolve_prep(R, *P);
            bar.wait();
            auto t0 = std::chrono::steady_clock::now();
            for (int s=0; s<nsub; ++s)
            {
                explicitSolve_substep(R, *P, dte);
                auto& vec = R.D("M_VT");
                for (int i=0; i<(int)R.M_extract_local_index.size(); i++)
                {
                    int const srl = (int)R.M_extract_local_index[i].size();
                    msg[r][i].resize(2*srl);
                    for (int j=0; j<srl; j++)
                    {
                        msg[r][i][j] = vec[R.M_extract_local_index[i][j]];
                        msg[r][i][j+srl] = vec[R.M_extract_local_index[i][j]+R.M_num_nodes];
                    }
                }
                bar.wait();
                for (int i=0; i<(int)R.M_local_ghosts_local_index.size(); i++)
                {
                    int const srl = (int)R.M_local_ghosts_local_index[i].size();
                    for (int j=0; j<srl; j++)
                    {
                        vec[R.M_local_ghosts_local_index[i][j]] = msg[i][r][j];
                        vec[R.M_local_ghosts_local_index[i][j]+R.M_num_nodes] = msg[i][r][j+srl];
                    }
                }
                bar.wait();
                if ( P->dynamics_type != 4 )
                    moveMesh(R, dte);
            }
            bar.wait();
            secs[r] = std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
        };
        std::vector<std::thread> th;
        for (int r=0;r<nranks;++r) th.emplace_back(body, r);
        for (auto& t : th) t.join();
        return *std::max_element(secs.begin(), secs.end());
    } catch (std::exception const& e) { g_err = e.what(); return -1.; }
}

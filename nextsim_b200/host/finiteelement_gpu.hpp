// finiteelement_gpu.hpp -- host-side C++ shim above the C ABI (include/nsx.h).
//
// Mirrors the slice of Nextsim::FiniteElement that the accelerated path replaces, with the reference's member
// and method names (model/finiteelement.hpp:166-169, 298-303, 521-522; model/finiteelement.cpp:8204-8212):
//
//     M_timer.tick("explicitSolve");  this->explicitSolve();  ...  this->update(UM_P);
//
// A neXtSIM maintainer keeps FiniteElement as it is and forwards these three member functions to an instance of
// FiniteElementGPU that points at the same std::vector<double> members.  Errors of the C
// ABI become the std::runtime_error the reference would throw (main.cpp:21-37 never catches them).
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/nsx.h"

namespace Nextsim {

class FiniteElementGPU
{
public:
    // what distributedMeshProcessing() (FE.cpp:50-143) leaves in FiniteElement, by reference
    struct MeshView {
        int M_num_nodes, M_local_ndof, M_num_elements, M_local_nelements;
        std::vector<double> const* coordX;              // M_mesh.coordX()
        std::vector<double> const* coordY;              // M_mesh.coordY()
        std::vector<int> const* indexTr;                // M_mesh.indexTr(), 1-based, 3 per element
        std::vector<unsigned char> const* mask_dirichlet;   // M_mask_dirichlet as bytes
        std::vector<int> const* M_neumann_flags;
        double const* NodalElementConnectivity; int nec_width;     // bamgmesh->NodalElementConnectivity[Size[1]]
        double const* NodalConnectivity; int nc_width;             // bamgmesh->NodalConnectivity[Size[1]]
        std::vector<double> const* lat;                 // M_mesh.lat()
        // initUpdateGhosts() products (FE.cpp:14003-14088); empty for one rank
        int rank, nranks;
        std::vector<int> const* M_recipients_proc_id;
        std::vector<std::vector<int>> const* M_extract_local_index;
        std::vector<int> const* M_local_ghosts_proc_id;
        std::vector<std::vector<int>> const* M_local_ghosts_local_index;
    };

    FiniteElementGPU(MeshView const& m, int device) { this->create(m, device); }
    ~FiniteElementGPU() { nsx_destroy(M_handle); }
    FiniteElementGPU(FiniteElementGPU const&) = delete;
    FiniteElementGPU& operator=(FiniteElementGPU const&) = delete;

    //! options: same keys and defaults as model/options.cpp (nsx_params_from_cfg reads a nextsim.cfg)
    void initOptAndParam(NsxDynParams const& p, double res_root_mesh)
    {
        M_params = p;
        // FE.cpp:6995-6998: scale_coef = sqrt(0.1/M_res_root_mesh); compr_strength *= scale_coef
        scale_coef = std::sqrt(0.1 / res_root_mesh);
        M_params.compr_strength = p.compr_strength * scale_coef;
        C_fix = p.C_lab * scale_coef;
        C_alea = p.alea_factor * C_fix;
        check(nsx_set_params(M_handle, &M_params), "nsx_set_params");
    }

    //! FiniteElement::calcCohesion (FE.cpp:3909-3914)
    void calcCohesion(std::vector<double> const& M_random_number, std::vector<double>& M_Cohesion) const
    {
        M_Cohesion.resize(M_random_number.size());
        for (size_t i = 0; i < M_random_number.size(); ++i) M_Cohesion[i] = C_fix + C_alea * M_random_number[i];
    }

    //! host -> device: every non-NULL member of `f` (NsxFields names are the FiniteElement member names)
    void upload(NsxFields const& f) { check(nsx_upload(M_handle, &f), "nsx_upload"); }
    //! device -> host
    void download(NsxFields& f) { check(nsx_download(M_handle, &f), "nsx_download"); }

    //! FiniteElement::explicitSolve() (FE.cpp:10182-10643)
    void explicitSolve() { check(nsx_explicit_solve(M_handle), "explicitSolve"); }
    //! FiniteElement::update(UM_P) (FE.cpp:3919-4132); UM_P is unused by the reference as well
    void update(std::vector<double> const& /*UM_P*/) { check(nsx_update(M_handle), "update"); }
    //! FiniteElement::updateGhosts(M_VT) (FE.cpp:13963-13996) on the device-resident velocity
    void updateGhosts() { check(nsx_update_ghosts(M_handle), "updateGhosts"); }

    //! FiniteElement::checkFieldsFast() (FE.cpp:14536-14655): throws like the reference when a field is off
    void checkFieldsFast()
    {
        NsxCheck c;
        check(nsx_check(M_handle, &c), "nsx_check");
        if (c.n_nan || c.n_speed || c.n_range)
            throw std::runtime_error("checkFieldsFast: " + std::to_string(c.n_nan) + " non-finite, " +
                                     std::to_string(c.n_speed) + " nodes above 5 m/s, " + std::to_string(c.n_range) +
                                     " elements out of range");
    }

    // ---- the callers either side of the path, on the resident state ----
    //! FiniteElement::checkRegridding() (FE.cpp:8298-8309).  `any_rank` stands for the reference's
    //! boost::mpi::all_reduce(M_comm, regrid_local, regrid, std::plus<bool>()) and stays with the host.
    template <class AllReduceOr>
    bool checkRegridding(double regrid_angle, AllReduceOr any_rank)
    {
        NsxRegrid r;
        check(nsx_check_regridding(M_handle, regrid_angle, &r), "checkRegridding");
        M_min_angle = r.min_angle;
        return any_rank(r.regrid != 0);
    }
    bool checkRegridding(double regrid_angle)
    {
        return this->checkRegridding(regrid_angle, [](bool b) { return b; });
    }
    //! FiniteElement::updateIceDiagnostics() (FE.cpp:7860-7900); results through download(D_conc, D_sigma, ...)
    void updateIceDiagnostics() { check(nsx_update_ice_diagnostics(M_handle), "updateIceDiagnostics"); }
    //! Dataset::variables[..].interpolated_data[slot] of M_wind / M_ocean / M_ssh after the spatial interpolation
    void loadForcing(int var, int slot, std::vector<double> const& interpolated_data)
    {
        check(nsx_forcing_load(M_handle, var, slot, interpolated_data.data()), "loadForcing");
    }
    //! ExternalData::getVector() (externaldata.cpp:366-455) evaluated into the resident M_wind / M_ocean / M_ssh
    void applyForcing(int var, bool interp_linear_time, double M_current_time, double ftime_range0, double ftime_range1,
                      double M_factor = 1., double M_bias_correction = 0.)
    {
        check(nsx_forcing_apply(M_handle, var, interp_linear_time ? 1 : 0, M_current_time, ftime_range0, ftime_range1,
                                M_factor, M_bias_correction), "applyForcing");
    }
    //! FiniteElement::thermo(int dt) (FE.cpp:5170-6137) on the resident state.  The caller keeps the options it read from
    //! `vm` in an NsxThermoParams (nsx_thermo_params_defaults() = model/options.cpp) and passes M_current_time.
    void thermo(NsxThermoParams const& p, int dt, double M_current_time)
    {
        check(nsx_thermo(M_handle, &p, dt, M_current_time), "thermo");
    }
    //! forcing / slab-ocean / tracer members of thermo() by the reference's member name ("M_tair", "M_sst", "D_Qa", ...)
    void thermoUpload(std::vector<std::pair<const char*, const double*>> const& fields)
    {
        std::vector<const char*> n; std::vector<const double*> v;
        for (auto const& f : fields) { n.push_back(f.first); v.push_back(f.second); }
        check(nsx_thermo_upload_many(M_handle, (int)n.size(), n.data(), v.data()), "thermoUpload");
    }
    void thermoDownload(std::vector<std::pair<const char*, double*>> const& fields)
    {
        std::vector<const char*> n; std::vector<double*> v;
        for (auto const& f : fields) { n.push_back(f.first); v.push_back(f.second); }
        check(nsx_thermo_download_many(M_handle, (int)n.size(), n.data(), v.data()), "thermoDownload");
    }
    double M_min_angle = 0.;    //!< "REGRID ANGLE" of the last checkRegridding (FE.cpp:8302)

    nsx_handle handle() const { return M_handle; }
    double scale_coef = 1., C_fix = 0., C_alea = 0.;

private:
    void check(int rc, const char* what) const
    {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + nsx_last_error(M_handle));
    }
    void create(MeshView const& m, int device)
    {
        NsxMesh M{};
        M.num_nodes = m.M_num_nodes; M.local_ndof = m.M_local_ndof;
        M.num_elements = m.M_num_elements; M.local_nelements = m.M_local_nelements;
        M.coord_x = m.coordX->data(); M.coord_y = m.coordY->data(); M.indices = m.indexTr->data();
        M.ghost_nodes = nullptr;
        M.mask_dirichlet = m.mask_dirichlet->data();
        M.neumann_flags = m.M_neumann_flags->data(); M.n_neumann_flags = (int)m.M_neumann_flags->size();
        M.nodal_element_connectivity = m.NodalElementConnectivity; M.nec_width = m.nec_width;
        M.nodal_connectivity = m.NodalConnectivity; M.nc_width = m.nc_width;
        M.lat = m.lat->data();
        NsxHalo H{};
        std::vector<int> sp, rp, sidx, ridx;
        if (m.nranks > 1) {
            H.rank = m.rank; H.nranks = m.nranks;
            sp.push_back(0); rp.push_back(0);
            for (int p : *m.M_recipients_proc_id) {
                auto const& v = (*m.M_extract_local_index)[p];
                sidx.insert(sidx.end(), v.begin(), v.end()); sp.push_back((int)sidx.size());
            }
            for (int p : *m.M_local_ghosts_proc_id) {
                auto const& v = (*m.M_local_ghosts_local_index)[p];
                ridx.insert(ridx.end(), v.begin(), v.end()); rp.push_back((int)ridx.size());
            }
            H.n_send_peers = (int)m.M_recipients_proc_id->size(); H.send_peer = m.M_recipients_proc_id->data();
            H.send_ptr = sp.data(); H.send_idx = sidx.data();
            H.n_recv_peers = (int)m.M_local_ghosts_proc_id->size(); H.recv_peer = m.M_local_ghosts_proc_id->data();
            H.recv_ptr = rp.data(); H.recv_idx = ridx.data();
        }
        if (nsx_create(&M, m.nranks > 1 ? &H : nullptr, device, &M_handle) != 0)
            throw std::runtime_error(std::string("nsx_create: ") + nsx_last_error(nullptr));
    }

    nsx_handle M_handle = nullptr;
    NsxDynParams M_params{};
};

}  // namespace Nextsim

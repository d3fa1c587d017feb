// host_step.cpp -- a C++ host process driving the path the way FiniteElement::step() does (FE.cpp:7963-8309),
// through the host library and the FiniteElementGPU shim only (no Python):
//
//   nextsim.cfg -> options            nsx_params_from_cfg + FiniteElementGPU::initOptAndParam   (FE.cpp:1047-1485, 6995-6999)
//   root mesh   -> local mesh, tables nsx_partmesh_build / _bc_marked_nodes / _views            (FE.cpp:50-143, 150-271)
//   cohesion                          FiniteElementGPU::calcCohesion                            (FE.cpp:3909-3914)
//   per step: upload -> explicitSolve() -> update() -> checkRegridding() -> updateIceDiagnostics() -> download
//
// usage: host_step <case.bin> <nextsim.cfg> <out.bin> <nsteps>
// case.bin (native endian): int nn, ne, n_dirichlet, n_neumann; double resolution; double x[nn], y[nn], lat[nn];
//   int tri[3*ne] (1-based); int dirichlet_flags_root[], neumann_flags_root[] (1-based); then the input fields as
//   doubles in the order of IN_FIELDS below (nodal vectors [u | v]).  out.bin: the output fields in OUT order, then
//   double min_angle, double regrid.
// Exit code 3 = nsx_create failed (no CUDA device): everything before it ran, nothing fell back to the CPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../nextsim_b200/host/finiteelement_gpu.hpp"

namespace {
struct Reader {
    FILE* f;
    template <class T> std::vector<T> vec(size_t n)
    {
        std::vector<T> v(n);
        if (n && std::fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "case file truncated\n"); std::exit(2); }
        return v;
    }
    template <class T> T one() { return vec<T>(1)[0]; }
};
void put(FILE* f, std::vector<double> const& v) { std::fwrite(v.data(), sizeof(double), v.size(), f); }
}

int main(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: host_step case.bin nextsim.cfg out.bin nsteps\n"); return 2; }
    FILE* fc = std::fopen(argv[1], "rb");
    if (!fc) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    Reader R{fc};
    int const nn = R.one<int>(), ne = R.one<int>(), nd = R.one<int>(), nm = R.one<int>();
    double const resolution = R.one<double>();
    auto x = R.vec<double>(nn), y = R.vec<double>(nn), lat = R.vec<double>(nn);
    auto tri = R.vec<int>(3 * (size_t)ne);
    auto dflags = R.vec<int>(nd), nflags = R.vec<int>(nm);

    // ---- options (same nextsim.cfg the reference reads)
    NsxDynParams P;
    if (nsx_params_from_cfg(argv[2], &P) != 0) { std::fprintf(stderr, "cfg: %s\n", nsx_cfg_last_error()); return 2; }

    // ---- mesh: one rank (the multi-rank wiring needs MPI)
    nsx_partmesh_handle pm = nullptr;
    if (nsx_partmesh_build(nn, x.data(), y.data(), ne, tri.data(), nullptr, nullptr, nullptr, 0, 1, &pm) != 0 ||
        nsx_partmesh_bc_marked_nodes(pm, dflags.data(), nd, nflags.data(), nm) != 0 || nsx_partmesh_set_lat(pm, lat.data()) != 0) {
        std::fprintf(stderr, "partmesh: %s\n", nsx_partmesh_last_error());
        return 2;
    }
    NsxMesh mesh{};
    NsxHalo halo{};
    nsx_partmesh_views(pm, &mesh, &halo);
    std::vector<double> coordX(mesh.coord_x, mesh.coord_x + nn), coordY(mesh.coord_y, mesh.coord_y + nn), mlat(mesh.lat, mesh.lat + nn);
    std::vector<int> indexTr(mesh.indices, mesh.indices + 3 * (size_t)ne), neumann(mesh.neumann_flags, mesh.neumann_flags + mesh.n_neumann_flags);
    std::vector<unsigned char> mask(mesh.mask_dirichlet, mesh.mask_dirichlet + nn);

    // ---- input fields
    auto nod2 = [&]() { return R.vec<double>(2 * (size_t)nn); };
    auto el = [&]() { return R.vec<double>((size_t)ne); };
    std::vector<double> M_VT = nod2(), M_UM = nod2(), M_UT = nod2(), M_wind = nod2(), M_ocean = nod2();
    std::vector<double> M_ssh = R.vec<double>(nn);
    std::vector<double> M_sigma[3] = {el(), el(), el()};
    std::vector<double> M_damage = el(), M_conc = el(), M_thick = el(), M_snow_thick = el(), M_conc_young = el(), M_h_young = el(),
                        M_hs_young = el(), M_thick_myi = el(), M_conc_myi = el(), M_ridge_ratio = el(), M_element_depth = el(),
                        M_drag_ui = el(), M_drag_ui_young = el(), M_random_number = el(), M_time_relaxation_damage = el();
    std::fclose(fc);
    std::printf("case: %d nodes, %d elements, %d dirichlet, %d neumann flags; nec width %d, nc width %d\n", nn, ne,
                (int)dflags.size(), mesh.n_neumann_flags, mesh.nec_width, mesh.nc_width);

    using Nextsim::FiniteElementGPU;
    FiniteElementGPU::MeshView mv{nn, mesh.local_ndof, ne, mesh.local_nelements, &coordX, &coordY, &indexTr, &mask, &neumann,
                                  mesh.nodal_element_connectivity, mesh.nec_width, mesh.nodal_connectivity, mesh.nc_width,
                                  &mlat, 0, 1, nullptr, nullptr, nullptr, nullptr};
    try {
        FiniteElementGPU* fe = nullptr;
        try {
            fe = new FiniteElementGPU(mv, 0);
        } catch (std::runtime_error const& e) {
            std::printf("threw: %s\n", e.what());
            return std::strstr(e.what(), "nsx_create") ? 3 : 1;
        }
        fe->initOptAndParam(P, resolution);
        std::vector<double> M_Cohesion;
        fe->calcCohesion(M_random_number, M_Cohesion);

        NsxFields in{};
        in.M_VT = M_VT.data(); in.M_UM = M_UM.data(); in.M_UT = M_UT.data(); in.M_wind = M_wind.data(); in.M_ocean = M_ocean.data();
        in.M_ssh = M_ssh.data();
        for (int k = 0; k < 3; ++k) in.M_sigma[k] = M_sigma[k].data();
        in.M_damage = M_damage.data(); in.M_conc = M_conc.data(); in.M_thick = M_thick.data(); in.M_snow_thick = M_snow_thick.data();
        in.M_conc_young = M_conc_young.data(); in.M_h_young = M_h_young.data(); in.M_hs_young = M_hs_young.data();
        in.M_thick_myi = M_thick_myi.data(); in.M_conc_myi = M_conc_myi.data(); in.M_ridge_ratio = M_ridge_ratio.data();
        in.M_element_depth = M_element_depth.data(); in.M_drag_ui = M_drag_ui.data(); in.M_drag_ui_young = M_drag_ui_young.data();
        in.M_Cohesion = M_Cohesion.data(); in.M_time_relaxation_damage = M_time_relaxation_damage.data();
        fe->upload(in);

        int const nsteps = std::atoi(argv[4]);
        bool regrid = false;
        for (int s = 0; s < nsteps; ++s) {
            fe->explicitSolve();                                   // FE.cpp:8204-8208
            fe->update(M_UM);                                      // FE.cpp:8210-8212
            fe->checkFieldsFast();
            regrid = fe->checkRegridding(10.);                     // numerics.regrid_angle default, FE.cpp:8298-8309
        }
        fe->updateIceDiagnostics();

        std::vector<double> D_tau_a(2 * (size_t)nn), D_tau_w(2 * (size_t)nn), M_surface(ne), D_conc(ne), D_thick(ne), D_snow(ne),
            D_s0(ne), D_s1(ne), D_div(ne);
        NsxFields out{};
        out.M_VT = M_VT.data(); out.M_UM = M_UM.data(); out.M_UT = M_UT.data(); out.D_tau_a = D_tau_a.data(); out.D_tau_w = D_tau_w.data();
        for (int k = 0; k < 3; ++k) out.M_sigma[k] = M_sigma[k].data();
        out.M_damage = M_damage.data(); out.M_conc = M_conc.data(); out.M_thick = M_thick.data(); out.M_snow_thick = M_snow_thick.data();
        out.M_ridge_ratio = M_ridge_ratio.data(); out.M_surface = M_surface.data();
        out.D_conc = D_conc.data(); out.D_thick = D_thick.data(); out.D_snow_thick = D_snow.data();
        out.D_sigma[0] = D_s0.data(); out.D_sigma[1] = D_s1.data(); out.D_divergence = D_div.data();
        fe->download(out);

        FILE* fo = std::fopen(argv[3], "wb");
        if (!fo) { std::fprintf(stderr, "cannot write %s\n", argv[3]); return 2; }
        put(fo, M_VT); put(fo, M_UM); put(fo, M_UT); put(fo, D_tau_a); put(fo, D_tau_w);
        for (int k = 0; k < 3; ++k) put(fo, M_sigma[k]);
        put(fo, M_damage); put(fo, M_conc); put(fo, M_thick); put(fo, M_snow_thick); put(fo, M_ridge_ratio); put(fo, M_surface);
        put(fo, D_conc); put(fo, D_thick); put(fo, D_snow); put(fo, D_s0); put(fo, D_s1); put(fo, D_div);
        put(fo, std::vector<double>{fe->M_min_angle, regrid ? 1. : 0.});
        std::fclose(fo);
        std::printf("ran %d steps, min angle %.6f, regrid %d\n", nsteps, fe->M_min_angle, (int)regrid);
        delete fe;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "host_step: %s\n", e.what());
        return 1;
    }
    nsx_partmesh_destroy(pm);
    return 0;
}

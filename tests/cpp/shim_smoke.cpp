// Compiles the host shim against include/nsx.h and libnsx.so and exercises the no-GPU behaviour: options and
// cfg parsing work on the host, creating a handle without a CUDA device throws (no CPU fallback).
#include <cstdio>
#include <cstring>
#include "../../nextsim_b200/host/finiteelement_gpu.hpp"

int main(int argc, char** argv)
{
    NsxDynParams p;
    nsx_params_defaults(&p);
    if (p.substeps != 120 || p.dtime_step != 200.) { std::printf("bad defaults\n"); return 1; }
    if (argc > 1 && nsx_params_from_cfg(argv[1], &p) != 0) { std::printf("cfg: %s\n", nsx_cfg_last_error()); return 1; }
    // one triangle
    std::vector<double> x{0., 1e4, 0.}, y{0., 0., 1e4}, lat{80., 80., 80.};
    std::vector<int> tr{1, 2, 3}, neu;
    std::vector<unsigned char> dir{0, 0, 0};
    double nec[3] = {1., 1., 1.};
    double nc[9] = {2., 3., 2., 1., 3., 2., 1., 2., 2.};
    Nextsim::FiniteElementGPU::MeshView m{3, 3, 1, 1, &x, &y, &tr, &dir, &neu, nec, 1, nc, 3, &lat, 0, 1, nullptr, nullptr, nullptr, nullptr};
    try {
        Nextsim::FiniteElementGPU fe(m, 0);
        fe.initOptAndParam(p, 7071.);
        std::printf("handle created (GPU present), scale_coef=%g\n", fe.scale_coef);
        // section 8(f) rows: instantiate the templates (reached only with a GPU)
        std::vector<double> w(6, 1.);
        fe.loadForcing(NSX_FORCING_WIND, 0, w);
        fe.applyForcing(NSX_FORCING_WIND, false, 0., 0., 1.);
        bool const regrid = fe.checkRegridding(10., [](bool b) { return b; }) || fe.checkRegridding(10.);
        std::printf("regrid=%d min angle=%g\n", (int)regrid, fe.M_min_angle);
        // row 3: thermo() on the one element (cold air over open water at the freezing point: new ice must form)
        NsxThermoParams tp;
        nsx_thermo_params_defaults(&tp);
        std::vector<double> tair{-20.}, dair{-22.}, mslp{101000.}, qsw{0.}, tcc{0.5}, precip{1e-5}, sst{-1.815}, sss{33.}, t0{-0.275};
        fe.thermoUpload({{"M_tair", tair.data()}, {"M_dair", dair.data()}, {"M_mslp", mslp.data()}, {"M_Qsw_in", qsw.data()},
                         {"M_tcc", tcc.data()}, {"M_precip", precip.data()}, {"M_sst", sst.data()}, {"M_sss", sss.data()},
                         {"M_tice0", t0.data()}, {"M_tice1", t0.data()}, {"M_tice2", t0.data()}, {"M_tsurf_young", t0.data()}});
        fe.thermo(tp, 200, 43133.25);
        std::vector<double> newice(1), hy(1);
        fe.thermoDownload({{"D_newice", newice.data()}, {"M_h_young", hy.data()}});
        std::printf("thermo: D_newice=%g m/day, M_h_young=%g m\n", newice[0], hy[0]);
        if (!(newice[0] > 0.) || !(hy[0] > 0.)) return 1;
    } catch (std::runtime_error const& e) {
        std::printf("threw: %s\n", e.what());
        return std::strstr(e.what(), "nsx_create") ? 0 : 1;
    }
    return 0;
}

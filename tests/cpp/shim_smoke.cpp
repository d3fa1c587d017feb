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
    } catch (std::runtime_error const& e) {
        std::printf("threw: %s\n", e.what());
        return std::strstr(e.what(), "nsx_create") ? 0 : 1;
    }
    return 0;
}

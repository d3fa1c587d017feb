// nsx_thermo.cu -- the thermo() kernel, in its own translation unit so that it can be compiled with -fmad=false:
// every product and sum then rounds exactly as in the reference's build (gcc without FMA contraction), and the device
// result differs from the CPU one only through exp / pow / log / cbrt / atan / hypot (<= 2 ulp each).  The kernel is bound
// by HBM traffic, so the unfused multiplies cost nothing measurable.
#include <cuda_runtime.h>

#include "nsx_thermo.cuh"

namespace nsx {

#ifndef NSX_THERMO_TPB
#define NSX_THERMO_TPB 128
#endif
#ifndef NSX_THERMO_MINB
#define NSX_THERMO_MINB 4
#endif
constexpr int THERMO_TPB = NSX_THERMO_TPB;

__global__ void __launch_bounds__(THERMO_TPB, NSX_THERMO_MINB)
k_thermo(const __grid_constant__ thermo::Params P, const __grid_constant__ thermo::Arrays A)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.ne) thermo::thermo_element(P, A, i);
}

cudaError_t launch_thermo(thermo::Params const& P, thermo::Arrays const& A, cudaStream_t stream)
{
    if (A.ne <= 0) return cudaSuccess;
    k_thermo<<<(A.ne + THERMO_TPB - 1) / THERMO_TPB, THERMO_TPB, 0, stream>>>(P, A);
    return cudaGetLastError();
}

}  // namespace nsx

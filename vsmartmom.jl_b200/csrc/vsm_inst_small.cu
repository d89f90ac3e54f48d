// small-matrix column kernel (vsm_small.cuh): NP = 16, four warps (= four spectral-point x moment items) per CTA
#include "vsm_small.cuh"

namespace vsm {
namespace small {
constexpr int kWPC = 4;
size_t small_smem_bytes() { return smem_bytes<16, kWPC>(); }
const void* small_kernel() { return (const void*)k_small<16, kWPC>; }
cudaError_t launch_small(cudaStream_t s, const ColArgs<double>& a, unsigned int total) {
    k_small<16, kWPC><<<(total + kWPC - 1) / kWPC, 32 * kWPC, smem_bytes<16, kWPC>(), s>>>(a, total);
    return cudaGetLastError();
}
}  // namespace small
}  // namespace vsm

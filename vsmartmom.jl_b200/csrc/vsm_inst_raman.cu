// rotational-Raman kernels (vsm_raman.cuh): NP = 16 (N <= 16: config 5, N = 15), NP = 32 (N <= 32) and NP = 48 (N <= 48)
#include "vsm_raman.cuh"

namespace vsm {
namespace rrs {

template <int NP, int WPC>
static cudaError_t le(cudaStream_t s, const Args& a, int m, int iz, int mode) {
    const int grid = (a.nE + WPC - 1) / WPC;
    k_rrs_elastic<NP, WPC><<<grid, 32 * WPC, elastic_smem<NP, WPC>(), s>>>(a, m, iz, mode);
    return cudaGetLastError();
}
template <int NP, int WPC>
static cudaError_t li(cudaStream_t s, const Args& a, int m, int iz, int mode, unsigned int total) {
    if (total == 0) return cudaSuccess;                    // no donor of this owner range lies inside the band
    const unsigned int grid = (total + WPC - 1) / WPC;
    k_rrs_inel<NP, WPC><<<grid, 32 * WPC, inel_smem<NP, WPC>(), s>>>(a, m, iz, mode, total);
    return cudaGetLastError();
}
template <int NP, int WPC>
static const Desc* make() {
    using G = Geo<NP>;
    static const Desc d{NP, WPC, elastic_smem<NP, WPC>(), inel_smem<NP, WPC>(),
                        (size_t)G::REC0, (size_t)G::REC1, (size_t)G::IREC0, (size_t)G::IREC1, (size_t)G::CE, (size_t)G::CI,
                        (const void*)k_rrs_elastic<NP, WPC>, (const void*)k_rrs_inel<NP, WPC>, le<NP, WPC>, li<NP, WPC>};
    return &d;
}
const Desc* desc(int N) {
    if (N <= 16) return make<16, 4>();
    if (N <= 32) return make<32, 4>();
    if (N <= 48) return make<48, 2>();
    return nullptr;
}
cudaError_t launch_fscatt(cudaStream_t s, int nE, int Nz, int nAer, size_t ldspec, const double* tau_rayl,
                          const double* aer_tau_mod, double* fsc) {
    k_rrs_fscatt<<<(nE + 127) / 128, 128, 0, s>>>(nE, Nz, nAer, ldspec, tau_rayl, aer_tau_mod, fsc);
    return cudaGetLastError();
}
cudaError_t launch_ss_rrs(cudaStream_t s, const SsArgs& a, int m, unsigned int nBand) {
    k_ss_rrs<<<(nBand + 3) / 4, 128, 0, s>>>(a, m, nBand);
    return cudaGetLastError();
}
cudaError_t launch_reduce(cudaStream_t s, const Args& a, double* ieR, double* ieT, int* nonfinite) {
    const size_t tot = (size_t)a.n1cnt * a.nVZA * a.nStokes;
    k_rrs_reduce<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(a, ieR, ieT, nonfinite);
    return cudaGetLastError();
}

}  // namespace rrs
}  // namespace vsm

// LinMode kernels on the generic (run-time N) Ops: any N, fp64
#include "vsm_launch.cuh"
#include "vsm_lin.cuh"
namespace vsm {
namespace {
template <int TM, int NTHR>
cudaError_t launch_layer_lin_g(cudaStream_t s, const ColArgs<double>& a, const LinArgs<double>& l, unsigned total, int grid, size_t smem) {
    k_layer_lin<double, OpsGeneric<double, TM>, 0, NTHR><<<grid, NTHR, smem, s>>>(a, l, total);
    return cudaGetLastError();
}
template <int TM, int NTHR>
cudaError_t launch_adding_lin_g(cudaStream_t s, const ColArgs<double>& a, const LinArgs<double>& l, unsigned total, int grid, size_t smem) {
    k_adding_lin<double, OpsGeneric<double, TM>, 0, NTHR><<<grid, NTHR, smem, s>>>(a, l, total);
    return cudaGetLastError();
}
const LinDesc kLin2 = {0, 0, 128, 0, (const void*)k_layer_lin<double, OpsGeneric<double, 2>, 0, 128>,
                       (const void*)k_adding_lin<double, OpsGeneric<double, 2>, 0, 128>, launch_layer_lin_g<2, 128>,
                       launch_adding_lin_g<2, 128>};
const LinDesc kLin4 = {0, 0, 256, 0, (const void*)k_layer_lin<double, OpsGeneric<double, 4>, 0, 256>,
                       (const void*)k_adding_lin<double, OpsGeneric<double, 4>, 0, 256>, launch_layer_lin_g<4, 256>,
                       launch_adding_lin_g<4, 256>};
}  // namespace
const LinDesc* lin_desc_generic(int tm) { return tm <= 2 ? &kLin2 : &kLin4; }
}  // namespace vsm

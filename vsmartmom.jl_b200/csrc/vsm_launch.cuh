// vsm_launch.cuh -- per-(precision, tile, size) kernel instantiations live in their
// own translation units (vsm_inst_*.cu) so the library builds in parallel.
//   NN = 0 : run-time N, generic arena (any size; spills to a global slab)
//   NN > 0 : compile-time N, one register tile per thread, all-shared-memory arena
#pragma once
#include "vsm_column.cuh"

namespace vsm {
template <typename T, int TM, int NN>
cudaError_t launch_layer(cudaStream_t s, const ColArgs<T>& a, unsigned total, int grid, int threads, size_t smem);
template <typename T, int TM, int NN>
cudaError_t launch_adding(cudaStream_t s, const ColArgs<T>& a, unsigned total, int grid, int threads, size_t smem);
template <typename T, int TM, int NN>
void kernel_ptrs(const void** layer, const void** adding);
template <typename T, int TM>
cudaError_t launch_step(cudaStream_t s, const StepArgs<T>& a, unsigned total, int grid, int threads, size_t smem);
template <typename T, int TM>
const void* step_kernel_ptr();
// the same step on the specialised Ops of the compile-time-N column kernels (L1 on the tensor-core path)
template <typename T, int NN>
cudaError_t launch_step_fast(cudaStream_t s, const StepArgs<T>& a, unsigned total, int grid, size_t smem);
template <typename T, int NN>
const void* step_fast_kernel_ptr();
// launch geometry chosen at compile time by the specialised instantiation
struct FastCfg {
    int layer_threads, adding_threads;
    int adding_two_group;      // 1: k_adding2 (11 matrices, two groups) ; 0: k_adding on OpsFast
};
template <typename T, int NN>
FastCfg fast_cfg();

// streamed-operand kernels (vsm_stream.cuh), fp64 only: one descriptor per instantiated N
struct StreamDesc {
    int N, LD, NP, MSZ;            // padded layout of every matrix (slab and hand-off): LD x NP elements
    int threads, layer_minb;
    size_t smem;
    const void* k_layer; const void* k_adding;
    cudaError_t (*launch_layer)(cudaStream_t, const ColArgs<double>&, unsigned total, int grid);
    cudaError_t (*launch_adding)(cudaStream_t, const ColArgs<double>&, unsigned total, int grid);
};
const StreamDesc* stream_desc_60();
const StreamDesc* stream_desc_75();
const StreamDesc* stream_desc_100();
inline const StreamDesc* stream_desc(int N) {
    return N == 60 ? stream_desc_60() : N == 75 ? stream_desc_75() : N == 100 ? stream_desc_100() : nullptr;
}

// LinMode kernels (vsm_lin.cuh), fp64: generic (run-time N, SIMT products on slab-resident matrices) and the
// streamed instantiations
template <typename T> struct LinArgs;
struct LinDesc {
    int ld, ncols;                 // slab / hand-off layout of one matrix: ld x ncols elements (0: N|1 x N, run-time)
    int threads;
    size_t smem;                   // 0: computed from N by the host (generic kernels)
    const void* k_layer; const void* k_adding;
    cudaError_t (*launch_layer)(cudaStream_t, const ColArgs<double>&, const LinArgs<double>&, unsigned total, int grid, size_t smem);
    cudaError_t (*launch_adding)(cudaStream_t, const ColArgs<double>&, const LinArgs<double>&, unsigned total, int grid, size_t smem);
};
const LinDesc* lin_desc_generic(int tm);        // tm = 2 (128 threads) or 4 (256 threads)
const LinDesc* lin_desc_60();
const LinDesc* lin_desc_75();
const LinDesc* lin_desc_100();
inline const LinDesc* lin_desc_stream(int N) {
    return N == 60 ? lin_desc_60() : N == 75 ? lin_desc_75() : N == 100 ? lin_desc_100() : nullptr;
}

// small-matrix column kernel (vsm_small.cuh): fp64, N <= 16, one warp per (point, moment), no hand-off
namespace small {
size_t small_smem_bytes();
const void* small_kernel();
cudaError_t launch_small(cudaStream_t, const ColArgs<double>&, unsigned int total);
}  // namespace small

template <typename T, int TM, int NN, int NTHR, int TMI = TM, int KS = 1>
struct OpsSel { using type = OpsFast<T, TM, NN, NTHR, 0, TMI, KS>; };
template <typename T, int TM, int NTHR, int TMI, int KS>
struct OpsSel<T, TM, 0, NTHR, TMI, KS> { using type = OpsGeneric<T, TM>; };
}  // namespace vsm

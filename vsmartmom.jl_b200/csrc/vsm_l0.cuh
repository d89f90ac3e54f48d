// vsm_l0.cuh -- batched dense algebra with the semantics of ext/gpu_batched_cuda.jl
// (batched_mul :208-218, batch_inv! :97-182, batch_solve! :72-94), one CTA per
// spectral slice, operands staged through shared memory when they fit.
#pragma once
#include "vsm_device.cuh"

namespace vsm {

// C(M x Nc) = A(M x K) * B(K x Nc), per batch element; column-major dense.
template <typename T, int TM, int TN>
__global__ void __launch_bounds__(256) k_bmm(T* __restrict__ C, const T* __restrict__ A,
                                             const T* __restrict__ B, int M, int K, int Nc,
                                             int batch, int stage) {
    T* sA = reinterpret_cast<T*>(vsm_smem);
    const int lda = M | 1, ldb = K | 1;
    T* sB = sA + (size_t)lda * K;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const T* Ab = A + (size_t)b * M * K;
        const T* Bb = B + (size_t)b * K * Nc;
        T* Cb = C + (size_t)b * M * Nc;
        const T* pa = Ab; const T* pb = Bb;
        int la = M, lb = K;
        if (stage) {
            __syncthreads();
            for (int e = threadIdx.x; e < M * K; e += blockDim.x) sA[(e % M) + (size_t)(e / M) * lda] = Ab[e];
            for (int e = threadIdx.x; e < K * Nc; e += blockDim.x) sB[(e % K) + (size_t)(e / K) * ldb] = Bb[e];
            __syncthreads();
            pa = sA; pb = sB; la = lda; lb = ldb;
        }
        const int nti = (M + TM - 1) / TM, ntj = (Nc + TN - 1) / TN;
        for (int tile = threadIdx.x; tile < nti * ntj; tile += blockDim.x) {
            const int i0 = (tile % nti) * TM, j0 = (tile / nti) * TN;
            int ia[TM], jb[TN];
#pragma unroll
            for (int x = 0; x < TM; ++x) ia[x] = min(i0 + x, M - 1);
#pragma unroll
            for (int y = 0; y < TN; ++y) jb[y] = min(j0 + y, Nc - 1) * lb;
            T acc[TM][TN];
#pragma unroll
            for (int x = 0; x < TM; ++x)
#pragma unroll
                for (int y = 0; y < TN; ++y) acc[x][y] = T(0);
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                T av[TM], bv[TN];
#pragma unroll
                for (int x = 0; x < TM; ++x) av[x] = pa[ia[x] + (size_t)k * la];
#pragma unroll
                for (int y = 0; y < TN; ++y) bv[y] = pb[jb[y] + k];
#pragma unroll
                for (int x = 0; x < TM; ++x)
#pragma unroll
                    for (int y = 0; y < TN; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
            }
#pragma unroll
            for (int y = 0; y < TN; ++y)
#pragma unroll
                for (int x = 0; x < TM; ++x)
                    if (i0 + x < M && j0 + y < Nc) Cb[(i0 + x) + (size_t)(j0 + y) * M] = acc[x][y];
        }
    }
}

// X = inv(A) per batch element.  stage=1: matrix in shared memory (odd ld);
// stage=0: in place in A (global), then copied to X.
template <typename T>
__global__ void __launch_bounds__(256) k_binv(T* __restrict__ X, T* __restrict__ A, int N, int batch,
                                              int* __restrict__ info, int stage) {
    T* sm = reinterpret_cast<T*>(vsm_smem);
    const int ld = stage ? (N | 1) : N;
    T* vec = stage ? sm + (size_t)ld * N : sm;
    int* piv = reinterpret_cast<int*>(vec + 3 * N);
    __shared__ int sing;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        T* Ab = A + (size_t)b * N * N;
        T* Xb = X + (size_t)b * N * N;
        T* Mx = stage ? sm : Ab;
        __syncthreads();
        if (threadIdx.x == 0) sing = 0;
        if (stage)
            for (int e = threadIdx.x; e < N * N; e += blockDim.x) Mx[(e % N) + (size_t)(e / N) * ld] = Ab[e];
        __syncthreads();
        cta_inverse<T>(N, Mx, ld, vec, piv, &sing);
        for (int e = threadIdx.x; e < N * N; e += blockDim.x) Xb[e] = Mx[(e % N) + (size_t)(e / N) * ld];
        if (info && threadIdx.x == 0) info[b] = sing;
    }
}

}  // namespace vsm

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

// ----------------------------------------------------------------------------
// fp64, N = 45 (the hot-path shape) fast paths: slices streamed through shared memory with 16-byte
// cp.async, product / inverse on the fp64 tensor cores (the OpsFast building blocks of the fused
// kernels).  Slices are only 8-byte aligned ((N*N)-strided), so the shared-memory copy is shifted
// by one element to the source's 16-byte phase; the odd element goes singly.
// ----------------------------------------------------------------------------
__device__ __forceinline__ double* cp_async_slice(double* dbase, const double* g, int count, int tid, int nthr) {
    const int ps = (int)((reinterpret_cast<uintptr_t>(g) >> 3) & 1);
    const int pd = (int)((dbase - reinterpret_cast<double*>(vsm_smem)) & 1);
    double* d = dbase + (ps ^ pd);
    const int head = ps, nvec = (count - head) / 2, tail0 = head + 2 * nvec;
    for (int v = tid; v < nvec; v += nthr) cp_async_16(d + head + 2 * v, g + head + 2 * v);
    if (tid == 0 && head) cp_async_elem(d, g);
    if (tid == 32 && tail0 < count) cp_async_elem(d + tail0, g + tail0);
    return d;
}

constexpr int kL0N = 45, kL0NN = kL0N * kL0N, kL0LB = kL0NN + 1;
using OpsL0 = OpsFast<double, 5, kL0N, 128, 0, 5, 1>;
static_assert((kL0N | 1) == kL0N, "dense slices need an odd N");

__global__ void __launch_bounds__(128, 3) k_bmm_f64_n45(double* __restrict__ C, const double* __restrict__ A,
                                                        const double* __restrict__ B, int batch) {
    double* sm = reinterpret_cast<double*>(vsm_smem);       // [2 stages][A, B], kL0LB each
    OpsL0 op;
    op.vec = nullptr; op.part = nullptr; op.piv = nullptr; op.dmask = nullptr; op.tid = threadIdx.x; op.bar = 0;
    const int tid = threadIdx.x;
    double *a0 = sm, *b0 = sm + kL0LB, *a1 = sm + 2 * kL0LB, *b1 = sm + 3 * kL0LB;
    int b = blockIdx.x;
    if (b < batch) {
        a0 = cp_async_slice(sm, A + (size_t)b * kL0NN, kL0NN, tid, 128);
        b0 = cp_async_slice(sm + kL0LB, B + (size_t)b * kL0NN, kL0NN, tid, 128);
    }
    cp_async_commit();
    for (int st = 0; b < batch; b += gridDim.x, st ^= 1) {
        const int nb = b + gridDim.x;
        if (nb < batch) {                                   // next slice into the other stage
            if (st) {
                a0 = cp_async_slice(sm, A + (size_t)nb * kL0NN, kL0NN, tid, 128);
                b0 = cp_async_slice(sm + kL0LB, B + (size_t)nb * kL0NN, kL0NN, tid, 128);
            } else {
                a1 = cp_async_slice(sm + 2 * kL0LB, A + (size_t)nb * kL0NN, kL0NN, tid, 128);
                b1 = cp_async_slice(sm + 3 * kL0LB, B + (size_t)nb * kL0NN, kL0NN, tid, 128);
            }
        }
        cp_async_commit();
        cp_async_wait_group<1>();                           // everything but the newest group has landed
        __syncthreads();
        double* Cb = C + (size_t)b * kL0NN;
        op.template mm_dmma_epi<false>(st ? a1 : a0, st ? b1 : b0, [&](int i, int j, double acc) { Cb[i + j * kL0N] = acc; });
        __syncthreads();                                    // this stage may be refilled two iterations on
    }
    cp_async_wait_all();
}

// X = inv(A): blocked Gauss-Jordan on the tensor cores with the threshold checks and the
// partial-pivoting fallback of OpsFast::inverse; info[b] = N if the result is not finite, else 0.
__global__ void __launch_bounds__(128, 3) k_binv_f64_n45(double* __restrict__ X, const double* __restrict__ A, int batch,
                                                         int* __restrict__ info) {
    double* sm = reinterpret_cast<double*>(vsm_smem);
    double* vec = sm + kL0LB;                               // 5n
    double* part = vec + 5 * kL0N;                          // 8n
    int* piv = reinterpret_cast<int*>(part + 8 * kL0N);     // n + 4
    OpsL0 op;
    op.vec = vec; op.part = part; op.piv = piv; op.dmask = nullptr; op.tid = threadIdx.x; op.bar = 0;
    const int tid = threadIdx.x;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * kL0NN;
        double* M = nullptr;
        auto load = [&]() {
            M = cp_async_slice(sm, Ab, kL0NN, tid, 128);
            cp_async_wait_all();
            __syncthreads();
        };
        __syncthreads();                                    // previous slice fully written out
        load();
        op.inverse(M, load);
        double* Xb = X + (size_t)b * kL0NN;
        int bad = 0;
        for (int e = tid; e < kL0NN; e += 128) {
            const double v = M[e];
            bad |= !(fabs(v) <= 1.7976931348623157e308);
            Xb[e] = v;
        }
        bad = __syncthreads_or(bad);
        if (info && tid == 0) info[b] = bad ? kL0N : 0;
    }
}

}  // namespace vsm

// vsm_device.cuh -- CTA-level building blocks for the fused adding-doubling kernels.
//
// One CTA owns the small (N x N) reflection / transmission matrices of one
// spectral point.  Matrices live in a per-CTA arena (shared memory first, a
// per-CTA global slab for whatever does not fit), column-major with an odd
// leading dimension so that both column reads (A operand) and strided row reads
// (B operand) are shared-memory bank-conflict free.
//
// Replaces, fused: NNlib.batched_mul / batch_inv! call chains of
//   src/CoreRT/CoreKernel/rt_helpers.jl:102-166 and interaction.jl:52-266
// (reference CUDA path: cuBLAS gemmStridedBatched + getrf/getriBatched,
//  ext/gpu_batched_cuda.jl:97-218).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vsm {

template <typename T> struct Num;
template <> struct Num<double> {
    __device__ static double eps() { return 2.220446049250313e-16; }
    __device__ static double exp_(double x) { return ::exp(x); }
    __device__ static double expm1_(double x) { return ::expm1(x); }
    __device__ static double abs_(double x) { return ::fabs(x); }
    // flip sign when mask != 0 (mask is 0 or 0x80000000): integer pipe, not FP64 pipe
    __device__ static double flip(double x, uint32_t mask) {
        int hi = __double2hiint(x) ^ (int)mask;
        return __hiloint2double(hi, __double2loint(x));
    }
};
template <> struct Num<float> {
    __device__ static float eps() { return 1.1920929e-07f; }
    __device__ static float exp_(float x) { return ::expf(x); }
    __device__ static float expm1_(float x) { return ::expm1f(x); }
    __device__ static float abs_(float x) { return ::fabsf(x); }
    __device__ static float flip(float x, uint32_t mask) {
        return __uint_as_float(__float_as_uint(x) ^ mask);
    }
};

// exp(-a) - exp(-b), stable for a ~ b   (rt_helpers.jl:32-40)
template <typename T>
__device__ inline T expdiff_neg(T a, T b) {
    if (a == b) return T(0);
    if (a < b) return Num<T>::exp_(-a) * (-Num<T>::expm1_(-(b - a)));
    return -Num<T>::exp_(-b) * (-Num<T>::expm1_(-(a - b)));
}

// D = diag(1,1,-1,-1) per stream: sign mask of Stokes component of row/col i
__device__ inline uint32_t dmask_of(int i, int nStokes) {
    return ((i % nStokes) >= 2) ? 0x80000000u : 0u;
}

// --------------------------------------------------------------------------
// C-tile GEMM:  acc(i,j) = sum_k A(i,k) * B(k,j),  then epi(i, j, acc).
// KSIGN: multiply the k-th term by D_k (folds r⁺⁻ = D r⁻⁺ D / t⁻⁻ = D t⁺⁺ D
// operands in without materialising them).  All threads of the CTA must call.
// Output must not alias A or B.
// --------------------------------------------------------------------------
template <typename T, int TM, int TN, bool KSIGN, class Epi>
__device__ __forceinline__ void cta_mm(int n, const T* __restrict__ A, int lda,
                                       const T* __restrict__ B, int ldb,
                                       const uint32_t* __restrict__ dmask, Epi epi) {
    const int nti = (n + TM - 1) / TM;
    const int ntj = (n + TN - 1) / TN;
    const int ntiles = nti * ntj;
    for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
        const int ti = tile % nti, tj = tile / nti;
        const int i0 = ti * TM, j0 = tj * TN;
        int ia[TM], jb[TN];
#pragma unroll
        for (int x = 0; x < TM; ++x) ia[x] = min(i0 + x, n - 1);
#pragma unroll
        for (int y = 0; y < TN; ++y) jb[y] = min(j0 + y, n - 1) * ldb;
        T acc[TM][TN];
#pragma unroll
        for (int x = 0; x < TM; ++x)
#pragma unroll
            for (int y = 0; y < TN; ++y) acc[x][y] = T(0);
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            T a[TM], b[TN];
            const T* Ak = A + (size_t)k * lda;
#pragma unroll
            for (int x = 0; x < TM; ++x) a[x] = Ak[ia[x]];
#pragma unroll
            for (int y = 0; y < TN; ++y) b[y] = B[jb[y] + k];
            if (KSIGN) {
                const uint32_t mk = dmask[k];
                if (TM <= TN) {
#pragma unroll
                    for (int x = 0; x < TM; ++x) a[x] = Num<T>::flip(a[x], mk);
                } else {
#pragma unroll
                    for (int y = 0; y < TN; ++y) b[y] = Num<T>::flip(b[y], mk);
                }
            }
#pragma unroll
            for (int x = 0; x < TM; ++x)
#pragma unroll
                for (int y = 0; y < TN; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
#pragma unroll
        for (int y = 0; y < TN; ++y)
#pragma unroll
            for (int x = 0; x < TM; ++x)
                if (i0 + x < n && j0 + y < n) epi(i0 + x, j0 + y, acc[x][y]);
    }
}

// --------------------------------------------------------------------------
// y(i) = sum_k A(i,k) x(k) [D_k], split over blockDim/nrow k-segments.
// part: scratch of size >= nseg*n.  Ends with the result handed to epi(i, y).
// Contains two __syncthreads(); x must be visible on entry.
// --------------------------------------------------------------------------
template <typename T, bool KSIGN, class Epi>
__device__ __forceinline__ void cta_mv(int n, const T* __restrict__ A, int lda,
                                       const T* __restrict__ x,
                                       const uint32_t* __restrict__ dmask,
                                       T* __restrict__ part, Epi epi) {
    const int nrow = (n + 31) & ~31;                 // rows padded to a warp multiple
    int nseg = 1;
    if ((int)blockDim.x >= nrow) {
        nseg = min((int)blockDim.x / nrow, 8);
        const int seg = threadIdx.x / nrow, i = threadIdx.x % nrow;
        const int klen = (n + nseg - 1) / nseg;
        if (seg < nseg && i < n) {
            const int k0 = seg * klen, k1 = min(n, k0 + klen);
            T s = T(0);
            for (int k = k0; k < k1; ++k) {
                T xv = x[k];
                if (KSIGN) xv = Num<T>::flip(xv, dmask[k]);
                s = fma(A[i + (size_t)k * lda], xv, s);
            }
            part[seg * n + i] = s;
        }
    } else {                                         // small CTAs: one segment, rows looped
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            T s = T(0);
            for (int k = 0; k < n; ++k) {
                T xv = x[k];
                if (KSIGN) xv = Num<T>::flip(xv, dmask[k]);
                s = fma(A[i + (size_t)k * lda], xv, s);
            }
            part[i] = s;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T s = part[i];
        for (int g = 1; g < nseg; ++g) s += part[g * n + i];
        epi(i, s);
    }
    __syncthreads();
}

// --------------------------------------------------------------------------
// In-place Gauss-Jordan inverse with partial (row) pivoting of an n x n matrix.
// Same pivot rule as LAPACK getrf (first row of maximum |a_ik|); the reference
// inverts with getrf+getri (CPU `A\I`, src/CoreRT/tools/cpu_batched.jl:32-37;
// cuBLAS getrf/getriBatched, ext/gpu_batched_cuda.jl:133-182).
// vec: scratch of 3n T;  piv: scratch of n int.  *sing_flag set to k+1 on a
// zero pivot.  Ends synchronised.
// --------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void cta_inverse(int n, T* __restrict__ M, int ld,
                                            T* __restrict__ vec, int* __restrict__ piv,
                                            int* sing_flag) {
    T* colb = vec;            // a_ik (post-swap)
    T* rowb = vec + n;        // a_kj / pivot (post-swap), rowb[k] = 1/pivot
    T* oldk = vec + 2 * n;    // row k before the swap
    const int tid = threadIdx.x;
    for (int k = 0; k < n; ++k) {
        if (tid < 32) {
            T best = T(-1);
            int bi = k;
            for (int i = k + tid; i < n; i += 32) {
                T v = Num<T>::abs_(M[i + (size_t)k * ld]);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                T ob = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) {
                piv[k] = bi;
                if (!(best > T(0)) && sing_flag) *sing_flag = k + 1;
            }
        }
        __syncthreads();
        const int p = piv[k];
        const T pv = M[p + (size_t)k * ld];
        const T ipv = T(1) / pv;
        for (int i = tid; i < n; i += blockDim.x) {
            // column k after the swap
            T c = (i == k) ? pv : ((i == p) ? M[k + (size_t)k * ld] : M[i + (size_t)k * ld]);
            colb[i] = c;
            // row p (-> new row k), scaled; and the old row k
            T rp = M[p + (size_t)i * ld];
            oldk[i] = M[k + (size_t)i * ld];
            rowb[i] = (i == k) ? ipv : rp * ipv;
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += blockDim.x) {
            const int i = e % n, j = e / n;
            T* mij = M + i + (size_t)j * ld;
            if (i == k) {
                *mij = rowb[j];
            } else {
                T base = (i == p) ? oldk[j] : *mij;
                if (j == k) base = T(0);
                *mij = fma(-colb[i], rowb[j], base);
            }
        }
        __syncthreads();
    }
    // undo the row permutation on the columns of the inverse
    for (int i = tid; i < n; i += blockDim.x) {
        for (int k = n - 1; k >= 0; --k) {
            const int p = piv[k];
            if (p != k) {
                T a = M[i + (size_t)k * ld], b = M[i + (size_t)p * ld];
                M[i + (size_t)k * ld] = b;
                M[i + (size_t)p * ld] = a;
            }
        }
    }
    __syncthreads();
}

}  // namespace vsm

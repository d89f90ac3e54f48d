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
    // branch-free reciprocal for |x| in the normal float range: fp32 seed + 2 Newton steps
    // in fp64 (23 -> 46 -> 92 bits): <= 1 ulp, no slow-path call on the critical path.
    __device__ static double rcp(double x) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));     // MUFU.RCP64H, ~20 bits
        double e = fma(-x, r, 1.0);
        r = fma(r, e, r);
        e = fma(-x, r, 1.0);
        return fma(r, e, r);
    }
    // |x| as an order-preserving integer key (high word), and that key scaled by 4
    __device__ static uint32_t key(double x) { return (uint32_t)__double2hiint(x) & 0x7fffffffu; }
    static constexpr uint32_t kKeyTimes4 = 0x00200000u;
    static constexpr uint32_t kKeyMin = 0x39b00000u, kKeyMax = 0x46200000u;   // ~1e-30 .. ~1e30
};
template <> struct Num<float> {
    __device__ static float eps() { return 1.1920929e-07f; }
    __device__ static float exp_(float x) { return ::expf(x); }
    __device__ static float expm1_(float x) { return ::expm1f(x); }
    __device__ static float abs_(float x) { return ::fabsf(x); }
    __device__ static float flip(float x, uint32_t mask) {
        return __uint_as_float(__float_as_uint(x) ^ mask);
    }
    __device__ static float rcp(float x) { return __frcp_rn(x); }
    __device__ static uint32_t key(float x) { return __float_as_uint(x) & 0x7fffffffu; }
    static constexpr uint32_t kKeyTimes4 = 0x01000000u;
    static constexpr uint32_t kKeyMin = 0x0d000000u, kKeyMax = 0x71000000u;   // ~4e-31 .. ~6e29
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

// ==========================================================================
// Ops policies.  The fused kernels are written against an Ops object providing
//   mm<KSIGN>(A, B, epi)                       epi(i, j, (A B)_ij)
//   mmv<NV>(A, B, epi, x1, x2, e1, e2)         ... plus y_v = A x_v as a by-product
//   mv<KSIGN>(A, x, epi), mv2(A, x1, x2, e1, e2)
//   inverse(M, rebuild)                        in place; rebuild() re-creates M for a pivoted retry
//   inv_one_minus_v<NV>(P, Q, X, S, x1, x2, e1, e2)    X = (I - P Q)^-1 (+ y_v = P x_v)
// OpsGeneric: run-time N, generic (shared-or-global) pointers, any size, SIMT.
// OpsFast   : compile-time N (NN), everything in shared memory; fp64: tensor-core (DMMA) products
//             and blocked Gauss-Jordan, adaptive Neumann inverse; fp32: one TMxTM register tile per
//             thread (SIMT) and the rank-1 register-resident Gauss-Jordan.
// ==========================================================================
template <typename T, int TM>
struct OpsGeneric {
    int n, ld;
    T* vec;                  // >= 5n
    T* part;                 // >= 8n
    int* piv;                // >= n+4
    const uint32_t* dmask;
    int tid, bar;            // whole CTA: tid = threadIdx.x
    int neumann = 0;         // (OpsFast only)
    static constexpr bool kDmma = false;
    static constexpr bool kIsStream = false;
    __device__ __forceinline__ int N() const { return n; }
    __device__ __forceinline__ int LD() const { return ld; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    template <bool KSIGN, class Epi>
    __device__ __forceinline__ void mm(const T* A, const T* B, Epi epi) const {
        cta_mm<T, TM, TM, KSIGN>(n, A, ld, B, ld, dmask, epi);
    }
    template <bool KSIGN, class Epi>
    __device__ __forceinline__ void mv(const T* A, const T* x, Epi epi) const {
        cta_mv<T, KSIGN>(n, A, ld, x, dmask, part, epi);
    }
    template <class E1, class E2>
    __device__ __forceinline__ void mv2(const T* A, const T* x1, const T* x2, E1 e1, E2 e2) const {
        cta_mv<T, false>(n, A, ld, x1, dmask, part, e1);
        cta_mv<T, false>(n, A, ld, x2, dmask, part, e2);
    }
    template <class Rebuild>
    __device__ __forceinline__ void inverse(T* M, Rebuild) const {
        cta_inverse<T>(n, M, ld, vec, piv, nullptr);
    }
    template <int NV, class Epi, class E1, class E2>
    __device__ __forceinline__ void mmv(const T* A, const T* B, Epi epi, const T* x1, const T* x2, E1 e1, E2 e2) const {
        cta_mm<T, TM, TM, false>(n, A, ld, B, ld, dmask, epi);
        if (NV >= 1) cta_mv<T, false>(n, A, ld, x1, dmask, part, e1);
        if (NV >= 2) cta_mv<T, false>(n, A, ld, x2, dmask, part, e2);
    }
    template <int NV, class E1, class E2>
    __device__ __forceinline__ void inv_one_minus_v(const T* P, const T* Q, T* X, T* S, const T* x1, const T* x2,
                                                    E1 e1, E2 e2) const {
        if (NV >= 1) cta_mv<T, false>(n, P, ld, x1, dmask, part, e1);
        if (NV >= 2) cta_mv<T, false>(n, P, ld, x2, dmask, part, e2);
        inv_one_minus(P, Q, X, S);
    }
    // X = (I - P Q)^-1 ; ends with the CTA synchronised (S: unused scratch)
    __device__ __forceinline__ void inv_one_minus(const T* P, const T* Q, T* X, T*) const {
        const int l = ld;
        cta_mm<T, TM, TM, false>(n, P, ld, Q, ld, dmask, [&](int i, int j, T acc) {
            X[i + j * l] = (i == j ? T(1) : T(0)) - acc;
        });
        __syncthreads();
        cta_inverse<T>(n, X, ld, vec, piv, nullptr);
    }
};

// The dynamic shared-memory arena of the fused kernels.  OpsFast re-bases every operand on
// this symbol (integer offset from it) so that the compiler addresses shared memory directly
// instead of converting generic pointers (a special-register read per access group).
extern __shared__ __align__(16) unsigned char vsm_smem[];
template <typename T>
__device__ __forceinline__ T* smem_base() { return reinterpret_cast<T*>(vsm_smem); }
template <typename T>
__device__ __forceinline__ T* rebase(const T* p) {
    return smem_base<T>() + (int)(p - smem_base<T>());
}

// One-element asynchronous global -> shared copy (LDGSTS; 4 or 8 bytes, natural alignment
// only, which is all the (N*N)-strided hand-off slots guarantee) and the matching wait.
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* dst_smem, const T* src_gmem) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "cp.async element size");
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(rebase(dst_smem));
    if (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src_gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src_gmem) : "memory");
}
template <typename T>
__device__ __forceinline__ void cp_async_16(T* dst_smem, const T* src_gmem) {     // both 16-byte aligned
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(rebase(dst_smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_gmem) : "memory");
}
// TMA bulk copy global -> shared (cp.async.bulk), completion counted on an mbarrier.  Source,
// destination and size are multiples of 16 bytes.  One thread issues; nothing blocks at issue.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}

#ifdef VSM_PHASE_TIMING
static __device__ unsigned long long vsm_inv_cycles[24];   // [0..7] CTA-wide callers, [8..15] named-barrier groups
#define VSM_NCOUNT(k) do { if (tid == 0) atomicAdd(&vsm_inv_cycles[16 + (k) + (BAR ? 4 : 0)], 1ull); } while (0)
#define VSM_IT0() long long _it0 = clock64()
#define VSM_ITICK(k)                                                             \
    do {                                                                         \
        long long _it1 = clock64();                                              \
        if (tid == 0) atomicAdd(&vsm_inv_cycles[(k) + (BAR ? 8 : 0)], (unsigned long long)(_it1 - _it0)); \
        _it0 = _it1;                                                             \
    } while (0)
#else
#define VSM_IT0()
#define VSM_ITICK(k)
#define VSM_NCOUNT(k)
#endif

// Named barrier over the NTHR threads of one Ops group (bar 0 == __syncthreads when
// the group is the whole CTA).
template <int NTHR, int BAR>
__device__ __forceinline__ void group_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NTHR) : "memory");
}

// TM: GEMM register tile, KS: split-K factor of the GEMMs (KS*(NN/TM)^2 threads work),
// TMI: register tile of the Gauss-Jordan inverse ((NN/TMI)^2 threads work).
template <typename T, int TM, int NN, int NTHR, int BAR = 0, int TMI = TM, int KS = 1>
struct OpsFast {
    static_assert(NN % TM == 0 && NN % TMI == 0, "OpsFast needs exact tiling");
    static constexpr int NT = NN / TM;            // GEMM tiles per dimension
    static constexpr int NTI = NN / TMI;          // inverse tiles per dimension
    static_assert(NT * NT * KS <= NTHR && NTI * NTI <= NTHR, "one register tile per thread");
    static constexpr int kLd = NN | 1;
    static constexpr bool kIsStream = false;
    static constexpr int RP = (NN + 31) & ~31;    // rows padded to whole warps for mv
    T* vec;                  // >= 5*NN : colb[2][NN] rowb[2][NN] ipv[2]
    T* part;                 // >= 8*NN
    int* piv;                // >= NN+4 ; piv[NN] = "unsafe pivot" flag
    const uint32_t* dmask;
    int tid;                 // thread index inside the group
    int bar;                 // (unused: the barrier id is the BAR template parameter)
    __device__ __forceinline__ constexpr int N() const { return NN; }
    __device__ __forceinline__ constexpr int LD() const { return kLd; }
    __device__ __forceinline__ void sync() const { group_sync<NTHR, BAR>(); }
    __device__ __forceinline__ int nthreads() const { return NTHR; }

    // SYNC_EPI: a group barrier between the last operand read and the epilogue, so that the
    // epilogue may overwrite A or B (each thread only touches the elements it owns).
    template <bool KSIGN, class Epi, bool SYNC_EPI = false>
    __device__ __forceinline__ void mm(const T* __restrict__ A, const T* __restrict__ B, Epi epi) const {
        if constexpr (sizeof(T) == 8 && (NTHR % 32 == 0) && (NTHR / 32 >= 4)) {
            mm_dmma_epi<KSIGN, Epi, SYNC_EPI>(A, B, epi);
            return;
        }
        const bool act = tid < NT * NT;
        const int i0 = (tid % NT) * TM, j0 = (tid / NT) * TM;
        T acc[TM][TM];
        if (act) {
            const T* Ap = rebase(A) + i0;
            const T* Bp = rebase(B) + j0 * kLd;
#pragma unroll
            for (int x = 0; x < TM; ++x)
#pragma unroll
                for (int y = 0; y < TM; ++y) acc[x][y] = T(0);
#pragma unroll 5
            for (int k = 0; k < NN; ++k) {
                T a[TM], b[TM];
#pragma unroll
                for (int x = 0; x < TM; ++x) a[x] = Ap[k * kLd + x];
#pragma unroll
                for (int y = 0; y < TM; ++y) b[y] = Bp[y * kLd + k];
                if (KSIGN) {
                    const uint32_t mk = dmask[k];
#pragma unroll
                    for (int x = 0; x < TM; ++x) a[x] = Num<T>::flip(a[x], mk);
                }
#pragma unroll
                for (int x = 0; x < TM; ++x)
#pragma unroll
                    for (int y = 0; y < TM; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
            }
        }
        if (SYNC_EPI) sync();
        if (act) {
#pragma unroll
            for (int y = 0; y < TM; ++y)
#pragma unroll
                for (int x = 0; x < TM; ++x) epi(i0 + x, j0 + y, acc[x][y]);
        }
    }

    // mm plus the side products y_v = A x_v (v < NV), see mm_dmma_v.  Ends WITHOUT a barrier on the
    // tensor-core path and with one on the SIMT path; callers synchronise before using any result.
    template <int NV, class Epi, class E1, class E2>
    __device__ __forceinline__ void mmv(const T* A, const T* B, Epi epi, const T* x1, const T* x2, E1 e1, E2 e2) const {
        if constexpr (kDmma) {
            mm_dmma_v<false, NV, false>(A, B, epi, x1, x2, e1, e2);
        } else {
            mm<false>(A, B, epi);
            if constexpr (NV > 0) mv_impl<false, NV>(A, x1, x2, e1, e2);
        }
    }

    // fp64 tensor-core GEMM (mma.sync m8n8k4): epi(i, j, (A*B)_ij).  Every lane of a
    // fragment load reads a distinct element, so shared-memory traffic per FMA is ~3x lower than
    // the broadcast-heavy register-tile SIMT product (which is shared-bandwidth-bound when all the
    // CTAs of an SM are in their GEMM phase).  fp64 DMMA shares the FMA datapath on B200 (same
    // 37 TFLOP/s, see profiles/microbench), so this is about operand bandwidth and issue slots.
    // The 45x45 operands are used unpadded: rows/columns >= NN read neighbouring data that only
    // lands in C rows/columns that are never stored; the K tail (k >= NN) is zeroed by predicate.
    static constexpr int kNW = NTHR / 32;
    static constexpr bool kDmma = (sizeof(T) == 8) && (NTHR % 32 == 0) && (kNW >= 4);
    // warp grid of the DMMA product: 2 x 2 (4 warps: 24x24 per warp) or 2 x 3 (>= 6 warps: 24x16)
    static constexpr int kWR = 2, kWC = (kNW >= 6) ? 3 : 2;
    // SYNC_EPI: a group barrier between the last operand read and the epilogue, so that the
    // epilogue may overwrite A or B (each thread only touches the elements it owns).
    // NV > 0: the product also yields y_v = A x_v for NV (1 or 2) vectors as one extra 8-column B
    // tile [x1 x2 0 ...] whose row tiles are dealt to the warps that already hold the A fragments;
    // e1(i, y1_i) / e2(i, y2_i) receive the result.  x1/x2 must not be written by e1/e2.
    template <bool KSIGN, class Epi, bool SYNC_EPI = false>
    __device__ __forceinline__ void mm_dmma_epi(const T* __restrict__ A, const T* __restrict__ B, Epi epi) const {
        auto none = [](int, T) {};
        mm_dmma_v<KSIGN, 0, SYNC_EPI>(A, B, epi, A, A, none, none);
    }
    template <bool KSIGN, int NV, bool SYNC_EPI, class Epi, class E1, class E2>
    __device__ __forceinline__ void mm_dmma_v(const T* __restrict__ A, const T* __restrict__ B, Epi epi,
                                              const T* __restrict__ x1, const T* __restrict__ x2, E1 e1, E2 e2) const {
        constexpr int NP = (NN + 7) & ~7;                   // padded extent (48)
        constexpr int TR = NP / 8 / kWR, TC = NP / 8 / kWC; // mma tiles per warp (rows, cols)
        constexpr int VS = (NV > 0) ? (TR + kWC - 1) / kWC : 1;   // vector-tile slots per warp
        // k tail: a last k-step with only 1 or 2 valid k (45 = 11*4 + 1) would spend a full DMMA per
        // tile on zeros; it is done as rank-1 DFMA updates on the accumulator fragments instead
        // (same datapath, 1/4 of the issue cycles per valid k).
        constexpr int KT = ((NN & 3) == 1 || (NN & 3) == 2) ? (NN & 3) : 0;
        constexpr int KM = KT ? (NN - KT) : ((NN + 3) & ~3);        // k range of the DMMA loop
        static_assert(NP / 8 % kWR == 0 && NP / 8 % kWC == 0, "warp tiling");
        const int warp = tid >> 5, lane = tid & 31;
        const int g = lane >> 2, tg = lane & 3;
        const int wr = warp % kWR, wc = warp / kWR;
        const int wi = wr * TR * 8, wj = wc * TC * 8;
        const bool wact = warp < kWR * kWC;
        double c[TR][TC][2];
        double cv[VS][2];
        if (wact) {
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int q = 0; q < TC; ++q) c[r][q][0] = c[r][q][1] = 0.0;
#pragma unroll
            for (int v = 0; v < VS; ++v) cv[v][0] = cv[v][1] = 0.0;
            // k of lane tg in DMMA step `st`: inside a block of 16 consecutive k the four lanes of a
            // fragment row take k = kb + s + 4*tg (s = 0..3 over four steps) instead of kb + 4s + tg.  Any
            // assignment works as long as A and B agree; with the odd leading dimension 45 this one
            // makes the 16 lanes of every half-warp hit 16 different 8-byte bank pairs for both
            // fragments (4*ld = 4 or 12 mod 16 for odd ld), where consecutive k gives 2-way conflicts on every load.
            constexpr int KB16 = (kLd & 1) ? (KM / 16) * 16 : 0;       // any odd leading dimension
            constexpr int NSTEP = (KM + 3) / 4;
            const T* Ar = rebase(A) + wi + g;                 // A[wi+g (+8r)][k]
            const T* Br = rebase(B) + (wj + g) * kLd;         // B[k][wj+g (+8c)]
            const T* xr = rebase((NV == 2 && g == 1) ? x2 : x1);
#pragma unroll
            for (int st = 0; st < NSTEP; ++st) {
                const int k0 = st * 4;
                const int k = (k0 < KB16) ? ((k0 & ~15) + ((k0 >> 2) & 3) + 4 * tg) : (k0 + tg);
                double a[TR], b[TC];
                const bool kin = (k0 + 3 < NN) || (k < NN);
                uint32_t mk = 0;
                if (KSIGN) mk = kin ? dmask[k] : 0u;
#pragma unroll
                for (int r = 0; r < TR; ++r) {
                    a[r] = kin ? (double)Ar[r * 8 + k * kLd] : 0.0;
                    if (KSIGN) a[r] = Num<double>::flip(a[r], mk);
                }
#pragma unroll
                for (int q = 0; q < TC; ++q) b[q] = kin ? (double)Br[k + q * 8 * kLd] : 0.0;
                double bx = 0.0;
                if (NV > 0) bx = (g < NV && kin) ? (double)xr[k] : 0.0;
#pragma unroll
                for (int r = 0; r < TR; ++r)
#pragma unroll
                    for (int q = 0; q < TC; ++q)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c[r][q][0]), "+d"(c[r][q][1]) : "d"(a[r]), "d"(b[q]));
                if (NV > 0) {
#pragma unroll
                    for (int v = 0; v < VS; ++v) {
                        double av = a[0];
#pragma unroll
                        for (int r = 1; r < TR; ++r)
                            if (r / kWC == v) av = (r % kWC == wc) ? a[r] : av;
                        if (v * kWC + wc < TR)                 // warp-uniform
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(cv[v][0]), "+d"(cv[v][1]) : "d"(av), "d"(bx));
                    }
                }
            }
            if (KT > 0) {
                // accumulator fragment: lane (g, tg) holds C[g][2tg], C[g][2tg+1] of each tile
                const T* At = rebase(A) + wi + g;             // A[wi + 8r + g][k]
                const T* Bt = rebase(B) + (wj + 2 * tg) * kLd;  // B[k][wj + 8q + 2tg + h]
#pragma unroll
                for (int k = KM; k < NN; ++k) {
                    double a[TR], b[TC][2];
#pragma unroll
                    for (int r = 0; r < TR; ++r) {
                        a[r] = (double)At[r * 8 + k * kLd];
                        if (KSIGN) a[r] = Num<double>::flip(a[r], dmask[k]);
                    }
#pragma unroll
                    for (int q = 0; q < TC; ++q) {
                        b[q][0] = (double)Bt[k + q * 8 * kLd];
                        b[q][1] = (double)Bt[k + (q * 8 + 1) * kLd];
                    }
#pragma unroll
                    for (int r = 0; r < TR; ++r)
#pragma unroll
                        for (int q = 0; q < TC; ++q) {
                            c[r][q][0] = fma(a[r], b[q][0], c[r][q][0]);
                            c[r][q][1] = fma(a[r], b[q][1], c[r][q][1]);
                        }
                    if (NV > 0) {
                        const double xk1 = (double)rebase(x1)[k];
                        const double xk2 = (NV == 2) ? (double)rebase(x2)[k] : 0.0;
#pragma unroll
                        for (int v = 0; v < VS; ++v) {
                            double av = a[0];
#pragma unroll
                            for (int r = 1; r < TR; ++r)
                                if (r / kWC == v) av = (r % kWC == wc) ? a[r] : av;
                            cv[v][0] = fma(av, xk1, cv[v][0]);       // only the tg == 0 lanes' values are used
                            cv[v][1] = fma(av, xk2, cv[v][1]);
                        }
                    }
                }
            }
        }
        if (SYNC_EPI) sync();
        if (wact) {
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int q = 0; q < TC; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = wi + r * 8 + g, j = wj + q * 8 + 2 * tg + h;
                        if (i < NN && j < NN) epi(i, j, (T)c[r][q][h]);
                    }
            if (NV > 0 && tg == 0) {
#pragma unroll
                for (int v = 0; v < VS; ++v) {
                    const int r = v * kWC + wc;
                    const int i = wi + r * 8 + g;
                    if (r < TR && i < NN) {
                        e1(i, (T)cv[v][0]);
                        if (NV == 2) e2(i, (T)cv[v][1]);
                    }
                }
            }
        }
    }
    // y = A x (one or two right-hand sides), k split over NTHR/RP segments
    template <bool KSIGN, int NV, class E1, class E2>
    __device__ __forceinline__ void mv_impl(const T* __restrict__ A, const T* __restrict__ x1,
                                            const T* __restrict__ x2, E1 e1, E2 e2) const {
        constexpr int NSEG = (NTHR / RP) < 1 ? 1 : (NTHR / RP);
        constexpr int KL = (NN + NSEG - 1) / NSEG;
        const int seg = tid / RP, i = tid % RP;
        if (seg < NSEG && i < NN) {
            const int k0 = seg * KL;
            T s1 = T(0), s2 = T(0);
#pragma unroll 4
            for (int kk = 0; kk < KL; ++kk) {
                const int k = k0 + kk;
                if (k < NN) {
                    const T av = A[i + k * kLd];
                    T xv = x1[k];
                    if (KSIGN) xv = Num<T>::flip(xv, dmask[k]);
                    s1 = fma(av, xv, s1);
                    if (NV == 2) s2 = fma(av, x2[k], s2);
                }
            }
            part[seg * NN + i] = s1;
            if (NV == 2) part[(NSEG + seg) * NN + i] = s2;
        }
        sync();
        constexpr bool SPLIT = (NTHR >= RP + NN);     // second vector reduced by a second thread range
        if (tid < NN) {
            T s = part[tid];
#pragma unroll
            for (int g = 1; g < NSEG; ++g) s += part[g * NN + tid];
            e1(tid, s);
            if (NV == 2 && !SPLIT) {
                T s2 = part[NSEG * NN + tid];
#pragma unroll
                for (int g = 1; g < NSEG; ++g) s2 += part[(NSEG + g) * NN + tid];
                e2(tid, s2);
            }
        } else if (NV == 2 && SPLIT && tid >= RP && tid < RP + NN) {
            const int r = tid - RP;
            T s = part[NSEG * NN + r];
#pragma unroll
            for (int g = 1; g < NSEG; ++g) s += part[(NSEG + g) * NN + r];
            e2(r, s);
        }
        sync();
    }
    // y = A x for one or two right-hand sides on the fp64 tensor cores: X = [x1 x2 0 ...] is the
    // 8-column B operand; warp w takes row tiles w, w+NWU, ...  Ends with the group synchronised.
    template <bool KSIGN, int NV, class E1, class E2>
    __device__ __forceinline__ void mv_dmma(const T* __restrict__ A, const T* __restrict__ x1,
                                            const T* __restrict__ x2, E1 e1, E2 e2) const {
        constexpr int NP = (NN + 7) & ~7, NB = NP / 8;
        constexpr int NWU = (kNW < NB) ? kNW : NB;
        constexpr int REP = (NB + NWU - 1) / NWU;
        const int warp = tid >> 5, lane = tid & 31;
        const int g = lane >> 2, tg = lane & 3;
        const T* Ab = rebase(A) + g + tg * kLd;
        const T* xs = rebase((NV == 2 && g == 1) ? x2 : x1);
        const bool xin = (g < NV);
        if (warp < NWU) {
            double c[REP][2];
#pragma unroll
            for (int r = 0; r < REP; ++r) c[r][0] = c[r][1] = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < NP; k0 += 4) {
                const bool kin = (k0 + 3 < NN) || (k0 + tg < NN);
                double b = (xin && kin) ? (double)xs[k0 + tg] : 0.0;
                if (KSIGN && kin) b = Num<double>::flip(b, dmask[k0 + tg]);
#pragma unroll
                for (int r = 0; r < REP; ++r) {
                    const int io = (warp + r * NWU) * 8;
                    if (io < NP) {
                        const double a = kin ? (double)Ab[io + k0 * kLd] : 0.0;
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c[r][0]), "+d"(c[r][1]) : "d"(a), "d"(b));
                    }
                }
            }
            if (tg == 0) {
#pragma unroll
                for (int r = 0; r < REP; ++r) {
                    const int i = (warp + r * NWU) * 8 + g;
                    if (i < NN) {
                        e1(i, (T)c[r][0]);
                        if (NV == 2) e2(i, (T)c[r][1]);
                    }
                }
            }
        }
        sync();
    }
    template <bool KSIGN, class Epi>
    __device__ __forceinline__ void mv(const T* A, const T* x, Epi epi) const {
        if constexpr (kDmma) mv_dmma<KSIGN, 1>(A, x, x, epi, epi);
        else mv_impl<KSIGN, 1>(A, x, x, epi, epi);
    }
    template <class E1, class E2>
    __device__ __forceinline__ void mv2(const T* A, const T* x1, const T* x2, E1 e1, E2 e2) const {
        if constexpr (kDmma) mv_dmma<false, 2>(A, x1, x2, e1, e2);
        else mv_impl<false, 2>(A, x1, x2, e1, e2);
    }

    // X = (I - P Q)^-1 with S as scratch; ends with the group synchronised.
    // A = P Q is formed first together with its Frobenius norm nu (a sub-multiplicative bound of
    // the spectral radius).  When nu is small the Neumann series
    //   (I - A)^-1 = (I + A)(I + A^2)(I + A^4)...      remainder  <= nu^K / (1 - nu)
    // is summed (fp64: on the tensor cores) up to the order K whose remainder is below 1e-20
    // (fp32: 1e-10) -- i.e. far below the rounding of the O(1) diagonal -- (K = 2: no product at all, K = 3: one, K = 4: two,
    // K = 8: four), replacing the latency-bound pivot chain of the elimination by dense products.
    // Thin layers (the vast majority of doubling and adding steps of a real atmosphere) take the
    // K = 2 exit.  Otherwise, or when `neumann` is 0, the blocked Gauss-Jordan inverse runs.
    int neumann = 1;
    __device__ __forceinline__ void inv_one_minus(const T* P, const T* Q, T* X, T* S) const {
        auto none = [](int, T) {};
        inv_one_minus_v<0>(P, Q, X, S, P, P, none, none);
    }
    // ... and y_v = P x_v as a by-product of forming P Q (see mmv)
    template <int NV, class E1, class E2>
    __device__ __forceinline__ void inv_one_minus_v(const T* P, const T* Q, T* X, T* S, const T* x1, const T* x2,
                                                    E1 e1, E2 e2) const {
        auto at = [](T* M, int i, int j) -> T& { return M[i + j * kLd]; };
        auto rebuild = [&]() {
            mm<false>(P, Q, [&](int i, int j, T acc) { at(X, i, j) = (i == j ? T(1) : T(0)) - acc; });
            sync();
        };
        if (neumann) {
            // remainder bound nu^K <= tol, a few hundred times below the rounding of the O(1) diagonal
            constexpr double tol = (sizeof(T) == 8) ? 1e-20 : 1e-10;
            constexpr T kNu2K2 = (T)tol;                                       // nu^2 <= tol
            constexpr T kNu2K3 = (T)((sizeof(T) == 8) ? 4.6e-14 : 2.1e-7);    // nu^3 <= tol
            constexpr T kNu2K4 = (T)((sizeof(T) == 8) ? 1e-10 : 1e-5);        // nu^4 <= tol
            constexpr T kNu2K8 = (T)((sizeof(T) == 8) ? 9e-6 : 3e-3);         // nu^8 <= tol
            T ss = T(0);
            mmv<NV>(P, Q, [&](int i, int j, T acc) { at(X, i, j) = acc; ss = fma(acc, acc, ss); }, x1, x2, e1, e2);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            T* np = part;
            if ((tid & 31) == 0) np[tid >> 5] = ss;
            sync();
            T nu2 = T(0);
#pragma unroll
            for (int w = 0; w < (NTHR + 31) / 32; ++w) nu2 += np[w];
            if (nu2 <= kNu2K2) {                                  // K = 2:  G = I + A
                VSM_NCOUNT(0);
                if (tid < NN) at(X, tid, tid) += T(1);
                sync();
                return;
            }
            if (nu2 <= kNu2K3) {                                  // K = 3:  G = I + A + A^2
                VSM_NCOUNT(1);
                auto e = [&](int i, int j, T acc) { at(X, i, j) = ((i == j ? T(1) : T(0)) + at(X, i, j)) + acc; };
                mm<false, decltype(e), true>(X, X, e);
                sync();
                return;
            }
            if (nu2 <= kNu2K4) {                                  // K = 4:  G = I + A + A^2 + A A^2
                VSM_NCOUNT(2);
                mm<false>(X, X, [&](int i, int j, T acc) { at(S, i, j) = acc; });
                sync();
                auto e = [&](int i, int j, T acc) { at(X, i, j) = ((i == j ? T(1) : T(0)) + at(X, i, j)) + (at(S, i, j) + acc); };
                mm<false, decltype(e), true>(X, S, e);
                sync();
                return;
            }
            if (nu2 <= kNu2K8) {                                  // K = 8:  ... (I + A^4)
                VSM_NCOUNT(3);
                mm<false>(X, X, [&](int i, int j, T acc) { at(S, i, j) = acc; });
                sync();
                auto e1 = [&](int i, int j, T acc) { at(X, i, j) = ((i == j ? T(1) : T(0)) + at(X, i, j)) + (at(S, i, j) + acc); };
                mm<false, decltype(e1), true>(X, S, e1);
                sync();
                auto e2 = [&](int i, int j, T acc) { at(S, i, j) = acc; };
                mm<false, decltype(e2), true>(S, S, e2);
                sync();
                auto e3 = [&](int i, int j, T acc) { at(X, i, j) += acc; };
                mm<false, decltype(e3), true>(X, S, e3);
                sync();
                return;
            }
            for (int e = tid; e < NN * NN; e += NTHR) {           // X = I - A, then eliminate
                const int i = e % NN, j = e / NN;
                at(X, i, j) = (i == j ? T(1) : T(0)) - at(X, i, j);
            }
            sync();
            inverse(X, rebuild);
            return;
        }
        mmv<NV>(P, Q, [&](int i, int j, T acc) { at(X, i, j) = (i == j ? T(1) : T(0)) - acc; }, x1, x2, e1, e2);
        sync();
        inverse(X, rebuild);
    }

    // In-place inverse.  Register-resident Gauss-Jordan on the diagonal pivots, one
    // barrier per step.  Step k is the rank-1 update  M <- M - c' r'^T  with
    //   c'_i = a_ik (i != k), c'_k = a_kk - 1 ;  r'_j = a_kj / a_kk (j != k), r'_k = 1 + 1/a_kk
    // which produces the in-place Gauss-Jordan result for every element (pivot row,
    // pivot column and the rest) without per-element selects.  Column k, row k and the
    // current diagonal are exchanged through double-buffered shared vectors; every thread
    // forms the NEXT pivot and its reciprocal itself (a_{k+1,k+1} - a_{k+1,k} a_{k,k+1}/a_kk),
    // so the reciprocal chain never waits on a store + barrier.  Every multiplier is checked
    // against the threshold-pivoting bound |a_ik| <= 4|a_kk|; if a step violates it (or a
    // pivot is zero/NaN/out of range) the matrix is rebuilt by `rebuild()` and inverted by
    // the partial-pivoting routine, so the result is always that of a stable elimination.
    template <class Rebuild>
    __device__ __forceinline__ void inverse(T* M, Rebuild rebuild) const {
        if constexpr (kDmma) inverse_blk(M, rebuild);
        else inverse_rank1(M, rebuild);
    }

    // ---- blocked in-place Gauss-Jordan inverse on the fp64 tensor cores -------------------
    // The matrix is distributed over the warps as register-resident 8x8 DMMA accumulator tiles
    // (warp grid kWR x kWC, TR x TC tiles each) for the whole inversion; its shared-memory buffer
    // is reused as double-buffered staging for the panels.  Block step K (fully unrolled):
    //   a. owners stage the OLD row panel X_j = M_Kj and column panel C_i = M_iK; the owner warp of
    //      M_KK inverts it with shuffles (8 rank-1 steps, no CTA barrier) and stages P.     [barrier]
    //   b. column owners: M_iK <- L_i = -C_i P (staged for everyone); row owners: M_Kj <- P X_j;
    //      M_KK <- P.                                                                       [barrier]
    //   c. trailing update of every other tile:  M_ij += L_i X_j   (= M_ij - M_iK P M_Kj).
    // 6 block steps x 2 barriers replace 45 barrier-separated rank-1 steps and all O(N^3) work is
    // m8n8k4 products.  Rows/cols >= NN are a virtual identity.  No inter-block pivoting: each pivot
    // inside a diagonal block is checked against |a_ik| <= 4|a_kk| and each multiplier against
    // |L_ab| <= 16; on a violation the matrix is rebuilt and inverted with partial pivoting.
    template <class Rebuild>
    __device__ __forceinline__ void inverse_blk(T* M, Rebuild rebuild) const {
        constexpr int NP = (NN + 7) & ~7, NB = NP / 8;
        constexpr int TR = NB / kWR, TC = NB / kWC;
        constexpr int RLD = 12, CLD = NP + 4, PLD = 12;            // conflict-free fragment strides
        constexpr int SET = RLD * NP + CLD * 8;                    // one staging set (row + column panel)
        static_assert(2 * SET <= NN * (NN | 1), "staging must fit in the matrix buffer");
        static_assert(2 * PLD * 8 <= 13 * NN, "P staging must fit in vec+part");
        double* Ms = reinterpret_cast<double*>(rebase(M));
        double* Pst = reinterpret_cast<double*>(rebase(vec));      // 2 x (8 x 8, ld PLD)
        int* bad = piv + NN;
        const int warp = tid >> 5, lane = tid & 31;
        const int g = lane >> 2, tg = lane & 3;
        const bool wact = warp < kWR * kWC;
        const int wr = warp % kWR, wc = warp / kWR;
        if (tid == 0) *bad = 0;
        bool mybad = false;
        VSM_IT0();
        auto mma = [](double& c0, double& c1, double a, double b) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
        };
        // ---- load the register tiles (virtual identity outside NN x NN)
        double c[TR][TC][2];
        if (wact) {
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int q = 0; q < TC; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = (wr * TR + r) * 8 + g, j = (wc * TC + q) * 8 + 2 * tg + h;
                        c[r][q][h] = (i < NN && j < NN) ? Ms[i + j * kLd] : ((i == j) ? 1.0 : 0.0);
                    }
        }
        sync();                                                    // everyone has read M before staging reuses it
        VSM_ITICK(0);
#pragma unroll
        for (int K = 0; K < NB; ++K) {
            double* rowb = Ms + (K & 1) * SET;                     // X: 8 x NP, ld RLD  (element [k][j] at k + j*RLD)
            double* colb = rowb + RLD * NP;                        // C/L: NP x 8, ld CLD (element [i][k] at i + k*CLD)
            double* Ps = Pst + (K & 1) * PLD * 8;                  // P: 8 x 8, ld PLD
            constexpr int KR = 0; (void)KR;
            const int kr = K / TR, kq = K / TC;                    // warp-grid coordinates of the pivot block
            const int r0 = K % TR, q0 = K % TC;                    // its position inside the owner warp
            // ---- a. stage old panels; invert the pivot block
            if (wact) {
                if (wr == kr) {                                    // I own tiles (K, j): stage X_j
#pragma unroll
                    for (int q = 0; q < TC; ++q) {
                        const int jo = (wc * TC + q) * 8;
                        rowb[g + (jo + 2 * tg) * RLD] = c[r0][q][0];
                        rowb[g + (jo + 2 * tg + 1) * RLD] = c[r0][q][1];
                    }
                }
                if (wc == kq) {                                    // I own tiles (i, K): stage C_i
#pragma unroll
                    for (int r = 0; r < TR; ++r) {
                        const int io = (wr * TR + r) * 8;
                        colb[io + g + (2 * tg) * CLD] = c[r][q0][0];
                        colb[io + g + (2 * tg + 1) * CLD] = c[r][q0][1];
                    }
                }
                if (wr == kr && wc == kq) {                        // pivot-block inverse (this warp only)
                    double d0 = c[r0][q0][0], d1 = c[r0][q0][1];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double x0 = __shfl_sync(0xffffffffu, d0, 4 * k + tg);          // D[k][2tg]
                        const double x1 = __shfl_sync(0xffffffffu, d1, 4 * k + tg);          // D[k][2tg+1]
                        const double cv = __shfl_sync(0xffffffffu, (k & 1) ? d1 : d0, 4 * g + (k >> 1));   // D[g][k]
                        const double pv = __shfl_sync(0xffffffffu, (k & 1) ? d1 : d0, 4 * k + (k >> 1));   // D[k][k]
                        const uint32_t kp = Num<double>::key(pv);
                        if (kp < Num<double>::kKeyMin || kp > Num<double>::kKeyMax) mybad = true;
                        if (g != k && Num<double>::key(cv) > kp + Num<double>::kKeyTimes4) mybad = true;
                        const double ipv = Num<double>::rcp(pv);
                        const double cc = (g == k) ? pv - 1.0 : cv;
                        const double y0 = (2 * tg == k) ? 1.0 + ipv : x0 * ipv;
                        const double y1 = (2 * tg + 1 == k) ? 1.0 + ipv : x1 * ipv;
                        d0 = fma(-cc, y0, d0);
                        d1 = fma(-cc, y1, d1);
                    }
                    c[r0][q0][0] = d0;                             // M_KK <- P
                    c[r0][q0][1] = d1;
                    Ps[g + (2 * tg) * PLD] = d0;
                    Ps[g + (2 * tg + 1) * PLD] = d1;
                }
            }
            VSM_ITICK(1);
            sync();
            VSM_ITICK(2);
            // ---- b. panels
            if (wact) {
                if (wc == kq) {                                    // L_i = -C_i P   (A = C_i, B = P)
                    const double pb0 = Ps[tg + g * PLD], pb1 = Ps[4 + tg + g * PLD];
#pragma unroll
                    for (int r = 0; r < TR; ++r) {
                        const int io = (wr * TR + r) * 8;
                        if (!(wr == kr && r == r0)) {
                            const double a0 = -colb[io + g + tg * CLD], a1 = -colb[io + g + (4 + tg) * CLD];
                            double l0 = 0.0, l1 = 0.0;
                            mma(l0, l1, a0, pb0);
                            mma(l0, l1, a1, pb1);
                            if (Num<double>::key(l0) > 0x40300000u || Num<double>::key(l1) > 0x40300000u) mybad = true;
                            c[r][q0][0] = l0;
                            c[r][q0][1] = l1;
                            __syncwarp();
                            colb[io + g + (2 * tg) * CLD] = l0;    // publish L_i over C_i (rows owned by this warp)
                            colb[io + g + (2 * tg + 1) * CLD] = l1;
                        }
                    }
                }
                if (wr == kr) {                                    // M_Kj <- P X_j  (A = P, B = X_j)
                    const double pa0 = Ps[g + tg * PLD], pa1 = Ps[g + (4 + tg) * PLD];
#pragma unroll
                    for (int q = 0; q < TC; ++q) {
                        const int jo = (wc * TC + q) * 8;
                        if (!(wc == kq && q == q0)) {
                            const double b0 = rowb[tg + (jo + g) * RLD], b1 = rowb[4 + tg + (jo + g) * RLD];
                            double n0 = 0.0, n1 = 0.0;
                            mma(n0, n1, pa0, b0);
                            mma(n0, n1, pa1, b1);
                            c[r0][q][0] = n0;
                            c[r0][q][1] = n1;
                        }
                    }
                }
            }
            VSM_ITICK(3);
            sync();
            VSM_ITICK(4);
            // ---- c. trailing update  M_ij += L_i X_j   (i != K, j != K)
            if (wact) {
                double la[TR][2], xb[TC][2];
#pragma unroll
                for (int r = 0; r < TR; ++r) {
                    const int io = (wr * TR + r) * 8;
                    la[r][0] = colb[io + g + tg * CLD];
                    la[r][1] = colb[io + g + (4 + tg) * CLD];
                }
#pragma unroll
                for (int q = 0; q < TC; ++q) {
                    const int jo = (wc * TC + q) * 8;
                    xb[q][0] = rowb[tg + (jo + g) * RLD];
                    xb[q][1] = rowb[4 + tg + (jo + g) * RLD];
                }
#pragma unroll
                for (int r = 0; r < TR; ++r)
#pragma unroll
                    for (int q = 0; q < TC; ++q) {
                        const bool skip = (wr == kr && r == r0) || (wc == kq && q == q0);
                        if (!skip) {
                            mma(c[r][q][0], c[r][q][1], la[r][0], xb[q][0]);
                            mma(c[r][q][0], c[r][q][1], la[r][1], xb[q][1]);
                        }
                    }
            }
            VSM_ITICK(5);
        }
        if (mybad) *bad = 1;
        sync();                                                    // staging no longer read
        if (wact) {
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int q = 0; q < TC; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = (wr * TR + r) * 8 + g, j = (wc * TC + q) * 8 + 2 * tg + h;
                        if (i < NN && j < NN) Ms[i + j * kLd] = c[r][q][h];
                    }
        }
        sync();
        if (*bad) {                       // uniform over the group
#ifdef VSM_PHASE_TIMING
            if (tid == 0) atomicAdd(&vsm_inv_cycles[7 + (BAR ? 8 : 0)], 1ull);
#endif
            rebuild();
            sync();
            group_inverse_pivoted(M);
        }
#ifdef VSM_PHASE_TIMING
        if (tid == 0) atomicAdd(&vsm_inv_cycles[6 + (BAR ? 8 : 0)], 1ull);
#endif
    }

    template <class Rebuild>
    __device__ __forceinline__ void inverse_rank1(T* M, Rebuild rebuild) const {
        T* colb = rebase(vec);           // [2][NN]
        T* rowb = colb + 2 * NN;         // [2][NN]
        T* diab = rebase(part);          // [2][NN]  (part is free during the inverse)
        M = rebase(M);
        int* bad = piv + NN;
        const bool act = tid < NTI * NTI;
        const int i0 = (tid % NTI) * TMI, j0 = (tid / NTI) * TMI;
        const bool isdiag = act && (i0 == j0);
        T m[TMI][TMI];
        if (tid == 0) *bad = 0;
        T ipv = Num<T>::rcp(M[0]);
        if (act) {
#pragma unroll
            for (int y = 0; y < TMI; ++y)
#pragma unroll
                for (int x = 0; x < TMI; ++x) m[x][y] = M[(i0 + x) + (j0 + y) * kLd];
            if (j0 == 0) {
#pragma unroll
                for (int x = 0; x < TMI; ++x) colb[i0 + x] = m[x][0];
            }
            if (i0 == 0) {           // row 0, pre-scaled by 1/a_00
#pragma unroll
                for (int y = 0; y < TMI; ++y) rowb[j0 + y] = m[0][y] * ipv;
            }
            if (isdiag) {
#pragma unroll
                for (int x = 0; x < TMI; ++x) diab[i0 + x] = m[x][x];
            }
        }
        sync();
        bool mybad = false;
        const int ti = tid % NTI, tj = tid / NTI;
        // steps k = K*TMI + d : d is unrolled so every register-tile index is static.
        // rowb holds row k ALREADY scaled by 1/a_kk (its owners know the reciprocal).
#pragma unroll 1
        for (int K = 0; K < NTI; ++K) {
            const bool tiK = act && (ti == K), tjK = act && (tj == K);
#pragma unroll
            for (int d = 0; d < TMI; ++d) {
                const int k = K * TMI + d, k1 = k + 1;
                const int cur = (k & 1) * NN, nxt = NN - cur;
                const int d1 = (d + 1 < TMI) ? d + 1 : 0;
                const int K1 = (d + 1 < TMI) ? K : K + 1;
                VSM_IT0();
                T ipv_next = T(0);
                if (k1 < NN) {         // next pivot and its reciprocal: a thread-local recurrence
                    const T pn = fma(-colb[cur + k1], rowb[cur + k1], diab[cur + k1]);
                    ipv_next = Num<T>::rcp(pn);
                }
                if (act) {
                    const T pv = colb[cur + k];
                    const uint32_t kp = Num<T>::key(pv);
                    const uint32_t klim = kp + Num<T>::kKeyTimes4;
                    if (kp < Num<T>::kKeyMin || kp > Num<T>::kKeyMax) mybad = true;
                    T c[TMI], r[TMI];
#pragma unroll
                    for (int x = 0; x < TMI; ++x) {
                        const T cv = colb[cur + i0 + x];
                        if (x == d) {
                            if (!tiK && Num<T>::key(cv) > klim) mybad = true;
                            c[x] = tiK ? pv - T(1) : cv;
                        } else {
                            if (Num<T>::key(cv) > klim) mybad = true;
                            c[x] = cv;
                        }
                    }
#pragma unroll
                    for (int y = 0; y < TMI; ++y) {
                        const T rv = rowb[cur + j0 + y];
                        r[y] = (y == d && tjK) ? T(1) + ipv : rv;
                    }
                    // the column / row / diagonal the NEXT step needs are formed and published first
#pragma unroll
                    for (int x = 0; x < TMI; ++x) m[x][d1] = fma(-c[x], r[d1], m[x][d1]);
#pragma unroll
                    for (int y = 0; y < TMI; ++y)
                        if (y != d1) m[d1][y] = fma(-c[d1], r[y], m[d1][y]);
                    VSM_ITICK(0);
                    if (tj == K1) {
#pragma unroll
                        for (int x = 0; x < TMI; ++x) colb[nxt + i0 + x] = m[x][d1];
                    }
                    if (ti == K1) {
#pragma unroll
                        for (int y = 0; y < TMI; ++y) rowb[nxt + j0 + y] = m[d1][y] * ipv_next;
                    }
#pragma unroll
                    for (int x = 0; x < TMI; ++x)
#pragma unroll
                        for (int y = 0; y < TMI; ++y)
                            if (x != d1 && y != d1) m[x][y] = fma(-c[x], r[y], m[x][y]);
                    if (isdiag) {
#pragma unroll
                        for (int x = 0; x < TMI; ++x) diab[nxt + i0 + x] = m[x][x];
                    }
                }
                ipv = ipv_next;
                VSM_ITICK(1);
                sync();
                VSM_ITICK(2);
            }
        }
        if (mybad) *bad = 1;
        if (act) {
#pragma unroll
            for (int y = 0; y < TMI; ++y)
#pragma unroll
                for (int x = 0; x < TMI; ++x) M[(i0 + x) + (j0 + y) * kLd] = m[x][y];
        }
        sync();
        if (*bad) {                       // uniform over the group
            rebuild();
            sync();
            group_inverse_pivoted(M);
        }
    }

    // partial-pivoting fallback on the group's threads (same algorithm as cta_inverse)
    __device__ __forceinline__ void group_inverse_pivoted(T* __restrict__ M) const {
        T* colb = vec; T* rowb = vec + NN; T* oldk = vec + 2 * NN;
        for (int k = 0; k < NN; ++k) {
            if (tid < 32) {
                T best = T(-1); int bi = k;
                for (int i = k + tid; i < NN; i += 32) {
                    T v = Num<T>::abs_(M[i + k * kLd]);
                    if (v > best) { best = v; bi = i; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    T ob = __shfl_xor_sync(0xffffffffu, best, off);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (tid == 0) piv[k] = bi;
            }
            sync();
            const int p = piv[k];
            const T pv = M[p + k * kLd];
            const T ipv = T(1) / pv;
            for (int i = tid; i < NN; i += NTHR) {
                colb[i] = (i == k) ? pv : ((i == p) ? M[k + k * kLd] : M[i + k * kLd]);
                const T rp = M[p + i * kLd];
                oldk[i] = M[k + i * kLd];
                rowb[i] = (i == k) ? ipv : rp * ipv;
            }
            sync();
            for (int e = tid; e < NN * NN; e += NTHR) {
                const int i = e % NN, j = e / NN;
                T* mij = M + i + j * kLd;
                if (i == k) {
                    *mij = rowb[j];
                } else {
                    T base = (i == p) ? oldk[j] : *mij;
                    if (j == k) base = T(0);
                    *mij = fma(-colb[i], rowb[j], base);
                }
            }
            sync();
        }
        for (int i = tid; i < NN; i += NTHR) {
            for (int k = NN - 1; k >= 0; --k) {
                const int p = piv[k];
                if (p != k) {
                    const T x = M[i + k * kLd], y = M[i + p * kLd];
                    M[i + k * kLd] = y;
                    M[i + p * kLd] = x;
                }
            }
        }
        sync();
    }
};

}  // namespace vsm

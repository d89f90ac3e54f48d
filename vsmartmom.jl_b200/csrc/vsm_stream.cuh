// vsm_stream.cuh -- streamed-operand kernels for matrix sizes whose working set does not fit one
// SM's shared memory (N = 60 / 75 / 100: BASELINE.json configs[2], ocean_coxmunk.yaml with 11 / 20
// streams, Stokes IQU / IQUV).
//
// 9 composite/scratch matrices of N = 100 in fp64 are 720 KB; one CTA has 227 KB.  Distributed shared
// memory is not the way out (DSMEM moves 17-21 B/cycle/SM, the L2 ~42 B/cycle/SM), so every matrix of a
// spectral point lives in a per-CTA slab in global memory that stays L2-resident (148 x ~1 MB), in a
// zero-padded layout (LD x NP, NP = 8*ceil(N/8), LD = NP rounded up to 4 or 12 mod 16), and every product
//   C = A B
// streams k-panels of both operands through an NST-stage shared-memory ring with 16-byte cp.async
// (LDGSTS), accumulating the whole C in DMMA register fragments (mma.sync.m8n8k4.f64) of a WR x WC warp
// grid.  The zero padding is an invariant (products of padded operands are padded; epilogues never write
// outside N x N), so the inner loops carry no predicates, and LD = 4 or 12 (mod 16) makes every fragment
// load of both staged panels bank-conflict free with consecutive k.  The ring runs ACROSS products: while
// product p drains its last panels, the first panels of product p+1 are already in flight when the caller
// announces p+1 and its operands were complete before p started; the step orders below are arranged so
// that this holds for all but one product of a doubling step and for every product of an adding step.
//
// Replaces, fused (as vsm_column.cuh): doubling_helper! (CoreKernel/doubling.jl:38-99, rt_helpers.jl:102-166)
// and interaction_helper!(::ScatteringInterface_11) (interaction.jl:207-266) on cuBLAS
// gemmStridedBatched / getrf+getriBatched (ext/gpu_batched_cuda.jl:97-218).
#pragma once
#include "vsm_column.cuh"

namespace vsm {

template <int NN_, int WR_, int WC_, int KP_ = 16, int NST_ = 3>
struct StreamCfg {
    static constexpr int NN = NN_, WR = WR_, WC = WC_, KP = KP_, NST = NST_;
    static constexpr int NT = (NN + 7) / 8;                 // 8x8 tiles per dimension
    static constexpr int NP = NT * 8;                       // padded extent
    static constexpr int pick_ld(int np) { return (np % 16 == 4 || np % 16 == 12) ? np : pick_ld(np + 1); }
    static constexpr int LD = pick_ld(NP);                  // leading dimension in global memory and of the A stage
    static constexpr int MSZ = LD * NP;                     // elements per matrix
    static constexpr int LDB = KP + 4;                      // leading dimension of the B stage (k fastest)
    static constexpr int A_STAGE = LD * KP, B_STAGE = LDB * NP, STAGE = A_STAGE + B_STAGE;
    static constexpr int NPAN = (NP + KP - 1) / KP;
    static constexpr int NTHR = WR * WC * 32;
    static constexpr int TR = (NT + WR - 1) / WR, TC = (NT + WC - 1) / WC;   // largest tile block of a warp
    static constexpr int VEC = 6 + 8 + 5 + 4 + 8;           // vector scratch in units of NP doubles (see kernels)
    static_assert(KP % 8 == 0 && (LDB % 16 == 4 || LDB % 16 == 12), "B stage must be conflict free");
    static_assert(NST * STAGE >= MSZ, "the elimination fallback stages one whole matrix in the ring");
    static constexpr size_t smem_bytes() {
        return ((size_t)NST * STAGE + (size_t)VEC * NP + kMaxComp) * sizeof(double) + ((size_t)2 * NP + 8) * 4;
    }
};

// cp.async ring state (uniform over the CTA)
struct StreamRing {
    unsigned committed = 0;          // groups committed so far; group G lives in stage G % NST
    const double* pfA = nullptr;     // operands of the announced next product ...
    const double* pfB = nullptr;
    unsigned pfG0 = 0;               // ... the group of its panel 0 and the panels already issued
    int pfN = 0;
};

// Position of a warp in the WR x WC warp grid.  A warp's DMMA work is (its row tiles) x (its column tiles) and the
// tile counts differ by one when the grid does not divide NT; warps w, w+4, w+8, ... share an SM sub-partition
// (one DMMA pipe), so the grid is laid over the warps such that every sub-partition gets a balanced mix.
template <int WR, int WC>
__device__ __forceinline__ void stream_warp_coords(int warp, int& wr, int& wc) {
    if (WR == 4 && WC == 4) { wc = warp >> 2; wr = (warp + wc) & 3; }          // each sub-partition: every row and column once
    else if (WR == 2 && WC == 4) { wr = warp >> 2; wc = wr ? 3 - (warp & 3) : (warp & 3); }   // large row half with small column
    else { wr = warp % WR; wc = warp / WR; }
}

template <class CF>
__device__ __forceinline__ void stream_issue_panel(double* stages, int tid, unsigned G, const double* __restrict__ A,
                                                   const double* __restrict__ B, int kp) {
    constexpr int NP = CF::NP, kLd = CF::LD, KP = CF::KP, NTHR = CF::NTHR;
    double* sA = stages + (G % CF::NST) * CF::STAGE;
    double* sB = sA + CF::A_STAGE;
    const int k0 = kp * KP;
    const int kc = (NP - k0 < KP) ? (NP - k0) : KP;               // multiple of 8
    const double* gA = A + (size_t)k0 * kLd;                      // kc contiguous columns of A
    for (int c = tid; c < kc * kLd / 2; c += NTHR) cp_async_16(sA + 2 * c, gA + 2 * c);
    // B panel: kc consecutive k (kc/2 16-byte chunks) of every column
    if (kc == KP) {
        constexpr int CPC = KP / 2;
        for (int c = tid; c < NP * CPC; c += NTHR) {
            const int j = c / CPC, q = c % CPC;
            cp_async_16(sB + j * CF::LDB + 2 * q, B + j * kLd + k0 + 2 * q);
        }
    } else {
        const int cpc = kc / 2;
        for (int c = tid; c < NP * cpc; c += NTHR) {
            const int j = c / cpc, q = c - j * cpc;
            cp_async_16(sB + j * CF::LDB + 2 * q, B + j * kLd + k0 + 2 * q);
        }
    }
    cp_async_commit();
}

template <class CF>
struct OpsStream {
    using T = double;
    static constexpr int NN = CF::NN, NP = CF::NP, kLd = CF::LD, NTHR = CF::NTHR;
    static constexpr int WR = CF::WR, WC = CF::WC, TR = CF::TR, TC = CF::TC, NST = CF::NST, KP = CF::KP;
    static constexpr bool kDmma = true;
    static constexpr bool kIsStream = true;
    using Cfg = CF;
    T* stages;               // NST * STAGE
    T* vec;                  // >= 5 NP (elimination fallback)
    T* part;                 // >= 8 NP
    int* piv;                // >= NP + 4
    const uint32_t* dmask;   // [NP], 0 beyond NN
    int tid, bar;
    int neumann = 1;
    mutable StreamRing ring;

    __device__ __forceinline__ constexpr int N() const { return NN; }
    __device__ __forceinline__ constexpr int LD() const { return kLd; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ int nthreads() const { return NTHR; }

    // first tile of warp-grid row wr / column wc (balanced split of NT tiles)
    static __device__ __forceinline__ int t0(int w, int W) { return (CF::NT * w) / W; }

    // nothing in flight, nothing announced (before the ring memory is used for anything else)
    __device__ __forceinline__ void drain() const {
        cp_async_wait_all();
        ring.pfN = 0; ring.pfA = ring.pfB = nullptr;
        __syncthreads();
    }

    // C = A B on the tensor cores, epi(i, j, C_ij) for i, j < NN.  A, B: padded matrices in global memory,
    // complete and visible (a CTA barrier after their last write) before the call -- and, for an announced
    // product, before the product during which it was announced started.
    // KSIGN: term k multiplied by D_k.  NV > 0: also y_v = A x_v (x in shared memory) -> e1 / e2, as one
    // extra B tile owned by warp-grid column 0.  nA/nB: the next product's operands (prefetch), or null.
    template <bool KSIGN, int NV, class Epi, class E1, class E2>
    __device__ __forceinline__ void mm_stream(const T* A, const T* B, Epi epi,
                                              const T* x1, const T* x2, E1 e1, E2 e2,
                                              const T* nA, const T* nB) const {
        const int warp = tid >> 5, lane = tid & 31;
        const int g = lane >> 2, tg = lane & 3;
        int wr, wc;
        stream_warp_coords<WR, WC>(warp, wr, wc);
        const int tr0 = t0(wr, WR), trn = t0(wr + 1, WR) - tr0;
        const int tc0 = t0(wc, WC), tcn = t0(wc + 1, WC) - tc0;
        const int wi = tr0 * 8, wj = tc0 * 8;
        const bool vown = (NV > 0) && (wc == 0);
        // Tile block of this warp = a core of TRB x TCB tiles that every warp owns (no predicate in its k loop) plus,
        // when the grid does not divide NT, one extra tile row and / or column: each extra is ONE warp-uniform branch
        // around a block of DMMAs per k step.  (A predicate around every single mma.sync costs ~10 instructions of
        // BSSY / WARPSYNC / ISETP per DMMA: ncu showed DMMA at 10 % of the instruction stream that way.)
        constexpr int TRB = CF::NT / WR, TCB = CF::NT / WC;
        constexpr bool RU = (CF::NT % WR) != 0, CU = (CF::NT % WC) != 0;
        const bool xr = RU && (trn > TRB), xc = CU && (tcn > TCB);
        StreamRing& rs = ring;
        unsigned g0;
        int issued;
        if (rs.pfN > 0 && rs.pfA == A && rs.pfB == B) {
            g0 = rs.pfG0; issued = rs.pfN;             // announced: its first panels are already in flight
        } else {
            if (rs.pfN > 0) cp_async_wait_all();       // an announcement that did not come true
            __syncthreads();                           // every stage is free
            g0 = rs.committed; issued = 0;
            for (; issued < NST - 1 && issued < CF::NPAN; ++issued, ++rs.committed)
                stream_issue_panel<CF>(stages, tid, rs.committed, A, B, issued);
        }
        rs.pfN = 0; rs.pfA = nA; rs.pfB = nB;
        double c[TR][TC][2];
        double cv[TR][2];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            cv[r][0] = cv[r][1] = 0.0;
#pragma unroll
            for (int q = 0; q < TC; ++q) c[r][q][0] = c[r][q][1] = 0.0;
        }
        const T* xr_ = (NV == 2 && g == 1) ? x2 : x1;
#pragma unroll 1
        for (int kp = 0; kp < CF::NPAN; ++kp) {
            const unsigned G = g0 + (unsigned)kp;
            {   // all of this thread's copies of group G have landed
                const int allow = (int)(rs.committed - 1u - G);
                if (allow <= 0) cp_async_wait_group<0>();
                else if (allow == 1) cp_async_wait_group<1>();
                else if (allow == 2) cp_async_wait_group<2>();
                else cp_async_wait_group<3>();
            }
            __syncthreads();                           // panel kp visible to all; everyone is done with group G - 1
            // top up the ring: stage (G - 1) % NST is free now -> groups up to G + NST - 1, this product's panels
            // first, then the announced product's
            while (rs.committed < G + (unsigned)NST) {
                if (issued < CF::NPAN) {
                    stream_issue_panel<CF>(stages, tid, rs.committed, A, B, issued);
                    ++issued;
                } else if (rs.pfA && rs.pfN < CF::NPAN) {
                    if (rs.pfN == 0) rs.pfG0 = rs.committed;
                    stream_issue_panel<CF>(stages, tid, rs.committed, rs.pfA, rs.pfB, rs.pfN);
                    ++rs.pfN;
                } else {
                    break;
                }
                ++rs.committed;
            }
            const T* sA = stages + (G % NST) * CF::STAGE;
            const T* sB = sA + CF::A_STAGE;
            const int k0 = kp * KP;
            constexpr int K4 = (NN + 3) & ~3;          // k >= K4 only meets zero padding
            const int kc = (K4 - k0 < KP) ? (K4 - k0) : KP;
            const T* Ar = rebase(sA) + wi + g + tg * kLd;
            const T* Br = rebase(sB) + tg + (wj + g) * CF::LDB;
#pragma unroll 2
            for (int ks = 0; ks < kc / 4; ++ks) {
                double a[TR], b[TC];
                uint32_t mk = 0;
                if (KSIGN) mk = dmask[k0 + ks * 4 + tg];
#pragma unroll
                for (int r = 0; r < TR; ++r) {
                    a[r] = (r < TRB || xr) ? Ar[r * 8 + ks * 4 * kLd] : 0.0;
                    if (KSIGN) a[r] = Num<double>::flip(a[r], mk);
                }
#pragma unroll
                for (int q = 0; q < TC; ++q) b[q] = (q < TCB || xc) ? Br[ks * 4 + q * 8 * CF::LDB] : 0.0;
                auto mma = [](double& c0, double& c1, double x, double y) {
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c0), "+d"(c1) : "d"(x), "d"(y));
                };
#pragma unroll
                for (int r = 0; r < TRB; ++r)
#pragma unroll
                    for (int q = 0; q < TCB; ++q) mma(c[r][q][0], c[r][q][1], a[r], b[q]);
                if (RU) {
                    if (xr) {
#pragma unroll
                        for (int q = 0; q < TCB; ++q) mma(c[TR - 1][q][0], c[TR - 1][q][1], a[TR - 1], b[q]);
                    }
                }
                if (CU) {
                    if (xc) {
#pragma unroll
                        for (int r = 0; r < TRB; ++r) mma(c[r][TC - 1][0], c[r][TC - 1][1], a[r], b[TC - 1]);
                        if (RU) {
                            if (xr) mma(c[TR - 1][TC - 1][0], c[TR - 1][TC - 1][1], a[TR - 1], b[TC - 1]);
                        }
                    }
                }
                if (NV > 0) {
                    if (vown) {
                        const int k = k0 + ks * 4 + tg;
                        const double bx = (g < NV && k < NN) ? xr_[k] : 0.0;
#pragma unroll
                        for (int r = 0; r < TRB; ++r) mma(cv[r][0], cv[r][1], a[r], bx);
                        if (RU) {
                            if (xr) mma(cv[TR - 1][0], cv[TR - 1][1], a[TR - 1], bx);
                        }
                    }
                }
            }
        }
        // (the panels of an announced product keep landing while the epilogue runs)
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int q = 0; q < TC; ++q)
                if (r < trn && q < tcn) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = wi + r * 8 + g, j = wj + q * 8 + 2 * tg + h;
                        if (i < NN && j < NN) epi(i, j, c[r][q][h]);
                    }
                }
        if (NV > 0) {
            if (vown && tg == 0) {
#pragma unroll
                for (int r = 0; r < TR; ++r) {
                    const int i = wi + r * 8 + g;
                    if (r < trn && i < NN) {
                        e1(i, cv[r][0]);
                        if (NV == 2) e2(i, cv[r][1]);
                    }
                }
            }
        }
    }

    // ---- the Ops interface of vsm_device.cuh (no announcements: used by the rare interfaces) ----
    template <bool KSIGN, class Epi, bool SYNC_EPI = false>
    __device__ __forceinline__ void mm(const T* A, const T* B, Epi epi) const {
        auto none = [](int, T) {};
        mm_stream<KSIGN, 0>(A, B, epi, A, A, none, none, nullptr, nullptr);
    }
    template <int NV, class Epi, class E1, class E2>
    __device__ __forceinline__ void mmv(const T* A, const T* B, Epi epi, const T* x1, const T* x2, E1 e1, E2 e2) const {
        mm_stream<false, NV>(A, B, epi, x1, x2, e1, e2, nullptr, nullptr);
    }
    // y = A x with A in global memory: row i by thread i % 128 of k-segment i / 128.  Ends synchronised.
    template <bool KSIGN, int NV, class E1, class E2>
    __device__ __forceinline__ void mv_impl(const T* __restrict__ A, const T* __restrict__ x1, const T* __restrict__ x2,
                                            E1 e1, E2 e2) const {
        constexpr int RP = (NN + 31) & ~31;
        constexpr int NSEG = (NTHR / RP) < 1 ? 1 : ((NTHR / RP) > 4 ? 4 : (NTHR / RP));
        constexpr int KL = (NN + NSEG - 1) / NSEG;
        static_assert(2 * NSEG * NN <= 8 * NP, "partials must fit in part");
        const int seg = tid / RP, i = tid % RP;
        if (seg < NSEG && i < NN) {
            const int kb = seg * KL;
            T s1 = 0.0, s2 = 0.0;
#pragma unroll 8
            for (int kk = 0; kk < KL; ++kk) {
                const int k = kb + kk;
                if (k < NN) {
                    const T av = A[i + (size_t)k * kLd];
                    T xv = x1[k];
                    if (KSIGN) xv = Num<T>::flip(xv, dmask[k]);
                    s1 = fma(av, xv, s1);
                    if (NV == 2) s2 = fma(av, x2[k], s2);
                }
            }
            part[seg * NN + i] = s1;
            if (NV == 2) part[(NSEG + seg) * NN + i] = s2;
        }
        __syncthreads();
        if (tid < NN) {
            T s = part[tid];
#pragma unroll
            for (int q = 1; q < NSEG; ++q) s += part[q * NN + tid];
            e1(tid, s);
            if (NV == 2) {
                T u = part[NSEG * NN + tid];
#pragma unroll
                for (int q = 1; q < NSEG; ++q) u += part[(NSEG + q) * NN + tid];
                e2(tid, u);
            }
        }
        __syncthreads();
    }
    template <bool KSIGN, class Epi>
    __device__ __forceinline__ void mv(const T* A, const T* x, Epi epi) const { mv_impl<KSIGN, 1>(A, x, x, epi, epi); }
    template <class E1, class E2>
    __device__ __forceinline__ void mv2(const T* A, const T* x1, const T* x2, E1 e1, E2 e2) const {
        mv_impl<false, 2>(A, x1, x2, e1, e2);
    }

    // In-place inverse of a matrix in global memory by the partial-pivoting elimination, staged whole in
    // the (drained) ring.  The fallback of the adaptive inverse and the route of inverse_mode 0.
    template <class Rebuild>
    __device__ __forceinline__ void inverse(T* M, Rebuild) const {
        drain();
        T* S = stages;
        for (int e = tid; e < CF::MSZ; e += NTHR) S[e] = M[e];
        __syncthreads();
        cta_inverse<T>(NN, S, kLd, vec, piv, nullptr);
        for (int e = tid; e < NN * NN; e += NTHR) {
            const int i = e % NN, j = e / NN;
            M[i + (size_t)j * kLd] = S[i + j * kLd];
        }
        __syncthreads();
    }

    // Frobenius norm^2 of what the epilogue just accumulated in ss (one partial per thread); ends synchronised
    __device__ __forceinline__ T block_sum(T ss) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        __syncthreads();                               // part may still be read by a previous mv
        if ((tid & 31) == 0) part[tid >> 5] = ss;
        __syncthreads();
        T s = 0.0;
#pragma unroll
        for (int w = 0; w < NTHR / 32; ++w) s += part[w];
        __syncthreads();
        return s;
    }
    // Neumann order for nu^2 = ||A||_F^2 (see OpsFast::inv_one_minus_v): 2, 3, 4, 8, or 0 = eliminate
    __device__ __forceinline__ int neumann_order(T nu2) const {
        if (!neumann) return 0;
        if (nu2 <= 1e-20) return 2;
        if (nu2 <= 4.6e-14) return 3;
        if (nu2 <= 1e-10) return 4;
        if (nu2 <= 9e-6) return 8;
        return 0;
    }
    // X holds A = P Q (raw) on entry, (I - A)^-1 on exit; S: scratch matrix; ends synchronised.
    // K = 2 touches only the diagonal, so a product announced before the call stays valid unless it reads X.
    __device__ __forceinline__ void finish_inverse(int K, T* X, T* S) const {
        auto at = [](T* M, int i, int j) -> T& { return M[i + (size_t)j * kLd]; };
        if (K == 2) {
            if (tid < NN) at(X, tid, tid) += 1.0;
            __syncthreads();
            return;
        }
        if (K == 3) {                                   // I + A + A^2
            mm<false>(X, X, [&](int i, int j, T acc) { at(X, i, j) = ((i == j ? 1.0 : 0.0) + at(X, i, j)) + acc; });
            __syncthreads();
            return;
        }
        if (K == 4 || K == 8) {
            mm<false>(X, X, [&](int i, int j, T acc) { at(S, i, j) = acc; });
            __syncthreads();
            mm<false>(X, S, [&](int i, int j, T acc) { at(X, i, j) = ((i == j ? 1.0 : 0.0) + at(X, i, j)) + (at(S, i, j) + acc); });
            __syncthreads();
            if (K == 8) {                               // ... (I + A^4)
                mm<false>(S, S, [&](int i, int j, T acc) { at(S, i, j) = acc; });      // in place: all panels are consumed before the epilogue
                __syncthreads();
                mm<false>(X, S, [&](int i, int j, T acc) { at(X, i, j) += acc; });
                __syncthreads();
            }
            return;
        }
        for (int e = tid; e < NN * NN; e += NTHR) {     // X = I - A, then eliminate
            const int i = e % NN, j = e / NN;
            at(X, i, j) = (i == j ? 1.0 : 0.0) - at(X, i, j);
        }
        __syncthreads();
        inverse(X, 0);
    }
    template <int NV, class E1, class E2>
    __device__ __forceinline__ void inv_one_minus_v(const T* P, const T* Q, T* X, T* S, const T* x1, const T* x2,
                                                    E1 e1, E2 e2) const {
        T ss = 0.0;
        mm_stream<false, NV>(P, Q, [&](int i, int j, T acc) { X[i + (size_t)j * kLd] = acc; ss = fma(acc, acc, ss); },
                             x1, x2, e1, e2, nullptr, nullptr);
        finish_inverse(neumann_order(block_sum(ss)), X, S);
    }
    __device__ __forceinline__ void inv_one_minus(const T* P, const T* Q, T* X, T* S) const {
        auto none = [](int, T) {};
        inv_one_minus_v<0>(P, Q, X, S, P, P, none, none);
    }
};

// ----------------------------------------------------------------------------
// nd doublings on slab-resident r, t (doubling_helper!, doubling.jl:38-99) with the ring kept busy:
//   P1  X1 = r r (+ r [j+ j1-])      announces P3          X1 <- (I - r r)^-1  (Neumann / elimination)
//   P3  W  = r t                     announces P2
//   P2  X2 = t X1                    (P4 needs X2: the one bubble of the step)
//   P4  r += X2 W (+ X2 [v1 v2])     announces P5
//   P5  X1 = X2 t ; swap(t, X1)      announces the next step's P1
// ----------------------------------------------------------------------------
template <class CF>
__device__ __forceinline__ void doubling_stream(const OpsStream<CF>& op, int nd, double expk, double*& r, double*& t,
                                                double*& X1, double* X2, double* W, double* jp, double* jm, double* j1p,
                                                double* j1m, double* v1, double* v2, double* sjp = nullptr, double* sjm = nullptr,
                                                double* sv1 = nullptr, double* sv2 = nullptr) {
    using T = double;
    constexpr int n = CF::NN, ld = CF::LD;
    auto at = [](T* M, int i, int j) -> T& { return M[i + (size_t)j * ld]; };
    auto attenuate = [&]() {
        for (int i = op.tid; i < n; i += CF::NTHR) { j1p[i] = jp[i] * expk; j1m[i] = jm[i] * expk; }
    };
    if (nd > 0) { attenuate(); __syncthreads(); }
    for (int it = 0; it < nd; ++it) {
        T ss = 0.0;
        op.template mm_stream<false, 2>(r, r, [&](int i, int j, T acc) { at(X1, i, j) = acc; ss = fma(acc, acc, ss); },
                                        jp, j1m, [&](int i, T s) { v1[i] = j1m[i] + s; }, [&](int i, T s) { v2[i] = jp[i] + s; },
                                        r, t);
        const int K = op.neumann_order(op.block_sum(ss));
        op.finish_inverse(K, X1, X2);                 // (K > 2 voids the announcement; P3 then starts cold)
        auto none = [](int, T) {};
        op.template mm_stream<false, 0>(r, t, [&](int i, int j, T acc) { at(W, i, j) = acc; }, r, r, none, none, t, X1);
        __syncthreads();
        op.template mm_stream<false, 0>(t, X1, [&](int i, int j, T acc) { at(X2, i, j) = acc; }, r, r, none, none, nullptr, nullptr);
        __syncthreads();
        if (sjp) {      // per-source slot with expk = 1 (ThermalEmission), see doubling_loop
            op.mv2(r, sjp, sjm, [&](int i, T s) { sv1[i] = sjm[i] + s; }, [&](int i, T s) { sv2[i] = sjp[i] + s; });
            op.mv2(X2, sv1, sv2, [&](int i, T s) { sjm[i] = sjm[i] + s; }, [&](int i, T s) { sjp[i] = sjp[i] + s; });
        }
        op.template mm_stream<false, 2>(X2, W, [&](int i, int j, T acc) { at(r, i, j) += acc; }, v1, v2,
                                        [&](int i, T s) { jm[i] = jm[i] + s; }, [&](int i, T s) { jp[i] = j1p[i] + s; },
                                        X2, t);
        __syncthreads();
        expk = expk * expk;
        const bool more = it + 1 < nd;
        op.template mm_stream<false, 0>(X2, t, [&](int i, int j, T acc) { at(X1, i, j) = acc; }, r, r, none, none,
                                        more ? r : nullptr, more ? r : nullptr);
        if (more) attenuate();
        __syncthreads();
        T* tmp = t; t = X1; X1 = tmp;
    }
}

constexpr int kStreamLayerMats = 5;
constexpr int kStreamAddMats = 11;      // RA RB TC TD X1a X2a X1b X2b X3b + a generated (r, t) for non-scattering layers

template <class CF>
struct StreamSmem {
    double* stages; double* vb;
    __device__ StreamSmem() {
        stages = reinterpret_cast<double*>(vsm_smem);
        vb = stages + CF::NST * CF::STAGE;
    }
};

// k_layer for the streamed path: elemental + doublings for one (point, layer, moment); the result goes to
// the hand-off in the padded layout (zero padding written out, so that k_adding_stream uses it in place).
template <class CF, int MINB>
__global__ void __launch_bounds__(CF::NTHR, MINB) k_layer_stream(ColArgs<double> a, unsigned int total) {
    using T = double;
    constexpr int n = CF::NN, NP = CF::NP, ld = CF::LD;
    StreamSmem<CF> sm;
    T* vb = sm.vb;
    T* jp = vb;             T* jm = jp + NP;    T* j1p = jm + NP;   T* j1m = j1p + NP;
    T* v1 = j1m + NP;       T* v2 = v1 + NP;    T* part = v2 + NP;  /* 8 NP */
    T* vec = part + 8 * NP; /* 5 NP */          T* rowtmp = vec + 5 * NP; /* 4 NP */
    T* cap = rowtmp + 4 * NP; /* 8 NP */        T* Wc = cap + 8 * NP;     /* kMaxComp */
    uint32_t* dmask = reinterpret_cast<uint32_t*>(Wc + kMaxComp);
    int* piv = reinterpret_cast<int*>(dmask + NP);
    __shared__ unsigned int sh_item;
    OpsStream<CF> op;
    op.stages = sm.stages; op.vec = vec; op.part = part; op.piv = piv; op.dmask = dmask; op.tid = threadIdx.x; op.bar = 0;
    op.neumann = a.neumann;
    T* slab = a.slab + (size_t)blockIdx.x * a.slab_stride;       // zero-initialised by the host
    for (int i = threadIdx.x; i < NP; i += blockDim.x) dmask[i] = (i < n) ? dmask_of(i, a.nStokes) : 0u;
    unsigned int nextv = 0;
    if (threadIdx.x == 0) nextv = atomicAdd(a.counter, 1u);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_item = nextv;
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= total) break;
        if (threadIdx.x == 0) nextv = atomicAdd(a.counter, 1u);
        const int p = item % a.nS;
        const unsigned int rest = item / a.nS;
        const int mi = rest % a.Mb;
        const int si = rest / a.Mb;
        const int iz = a.layer_order[si];
        const int m = a.m0 + mi;
        const int nd = a.ndoubl[iz];
        const size_t pi = (size_t)iz * a.nS + p;
        const T tau = a.p_tau[pi], varpi = a.p_varpi[pi], tausum = a.p_tausum[pi];
        const T dtau = tau / (T)exp2((double)nd);
        const T expk = Num<T>::exp_(-dtau / a.mu0);
        if (threadIdx.x < a.nC) Wc[threadIdx.x] = a.p_W[((size_t)threadIdx.x * a.Nz + iz) * a.nS + p];
        __syncthreads();
        T* r = slab; T* t = slab + CF::MSZ; T* X1 = slab + 2 * CF::MSZ; T* X2 = slab + 3 * CF::MSZ; T* W = slab + 4 * CF::MSZ;
        const size_t zstride = (size_t)n * n;
        const T* Zpp = a.Zpp + (size_t)(m - a.mz0) * a.nC * zstride;
        const T* Zmp = a.Zmp + (size_t)(m - a.mz0) * a.nC * zstride;
        elemental_fill<T, CF::NN, CF::LD>(a, m, nd, dtau, varpi, tausum, Wc, Zpp, Zmp, zstride,
                                          a.F0 + (a.F0_off + p) * a.nStokes, dmask, r, t, jp, jm, rowtmp, cap);
        __syncthreads();
        const bool thermal = (a.therm_B != nullptr) && (m == 0);
        T* sjp = cap; T* sjm = cap + NP; T* sv1 = cap + 2 * NP; T* sv2 = cap + 3 * NP;       // (cap is free after the elemental phase)
        if (thermal) { thermal_fill<T>(a, n, iz, p, dtau, varpi, sjp, sjm); __syncthreads(); }
        doubling_stream<CF>(op, nd, expk, r, t, X1, X2, W, jp, jm, j1p, j1m, v1, v2, thermal ? sjp : nullptr, sjm, sv1, sv2);
        // apply_D_matrix! / apply_D_matrix_SFI! (doubling.jl:178-216) on the way out
        const size_t slot = ((size_t)mi * a.n_sched_layers + si) * a.nS + p;
        T* gr = a.add_r + slot * a.add_mstride;
        T* gt = a.add_t + slot * a.add_mstride;
        for (int e = threadIdx.x; e < CF::MSZ; e += blockDim.x) {
            const int i = e % ld, j = e / ld;
            const bool in = (i < n && j < n);
            T rv = in ? r[e] : 0.0;
            if (nd >= 1 && in) rv = Num<T>::flip(rv, dmask[i]);
            gr[e] = rv;
            gt[e] = in ? t[e] : 0.0;
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            T vm = jm[i], vp = jp[i];
            if (thermal) { vm += sjm[i]; vp += sjp[i]; }
            if (nd >= 1) vm = Num<T>::flip(vm, dmask[i]);
            a.add_jp[slot * n + i] = vp;
            a.add_jm[slot * n + i] = vm;
        }
    }
}

// One ScatteringInterface_11 step (interaction.jl:220-265) on slab-resident matrices; r, t: the added layer
// (read only: the hand-off slot itself, or the surface / generated layer).  The minus half (M*) and the plus
// half (P*) both read the pre-step composite; they are interleaved so that no product reads what the product
// before it wrote, and ordered so that R+-' and T--' exist two products before the next step needs them:
//   M1 X1a = r R+- (+ r J+)   P1 X1b = R+- r (+ R+- j-)   -> both inverses (Neumann K = 2: diagonal only)
//   M2 X2a = T-- X1a          P2 X2b = t X1b  (T21)
//   M3 X1a = X2a r (+ J-)     P4 X3b = T21 R+-
//   M4 R-+ += X1a T++         P5 R+-' = r+- + (X3b D t) D
//   P3 T++' = T21 T++ (+ J+)  M5 T--' = (X2a D t) D
template <class CF>
struct StreamComp {
    double *RA, *RB, *TC, *TD, *X1a, *X2a, *X1b, *X2b, *X3b;
    double *J0pc, *J0pnew, *J0m, *v1A, *v1B;
};

template <class CF>
__device__ __forceinline__ void add_step11_stream(const OpsStream<CF>& op, StreamComp<CF>& c, const double* r, const double* t,
                                                  const double* jp, const double* jm, bool rpm_zero, bool t_ident,
                                                  const double* next_r, bool full) {
    using T = double;
    constexpr int n = CF::NN, ld = CF::LD;
    const uint32_t* dmask = op.dmask;
    auto at = [](T* M, int i, int j) -> T& { return M[i + (size_t)j * ld]; };
    auto rd = [](const T* M, int i, int j) -> T { return M[i + (size_t)j * ld]; };
    auto none = [](int, T) {};
    T* J0m = c.J0m; T* v1A = c.v1A; T* v1B = c.v1B; T* J0pc = c.J0pc; T* J0pn = c.J0pnew;
    T ssa = 0.0, ssb = 0.0;
    if (!full) {
        // The composite R-+ and T++ enter no J0+- / T-- / R+- of this or any later step: without the rt_run_streams
        // export M4 and the matrix halves of M3 / P3 are dropped and the two source lines ride on M5 / P4:
        //   P1 X1b = R+- r (+ R+- j-)      M1 X1a = r R+- (+ r J+)       -> both inverses
        //   P2 X2b = t X1b (T21)           M2 X2a = T-- X1a
        //   P4 X3b = T21 R+- (+ J+')       M5 T--' = (X2a D t) D (+ J-)  P5 R+-' = r+- + (X3b D t) D
        // Every product but the step's first is announced by the one before it (its operands are complete by then).
        op.template mm_stream<false, 1>(c.RB, r, [&](int i, int j, T acc) { at(c.X1b, i, j) = acc; ssb = fma(acc, acc, ssb); },
                                        jm, jm, [&](int i, T s) { v1B[i] = J0pc[i] + s; }, none, r, c.RB);
        op.finish_inverse(op.neumann_order(op.block_sum(ssb)), c.X1b, c.X2b);
        // v1A is kept D-flipped: the source line of M5 runs through its D-signed k loop
        op.template mm_stream<false, 1>(r, c.RB, [&](int i, int j, T acc) { at(c.X1a, i, j) = acc; ssa = fma(acc, acc, ssa); },
                                        J0pc, J0pc, [&](int i, T s) { v1A[i] = Num<T>::flip(s + jm[i], dmask[i]); }, none,
                                        t_ident ? nullptr : t, t_ident ? nullptr : c.X1b);
        op.finish_inverse(op.neumann_order(op.block_sum(ssa)), c.X1a, c.X2a);
        const T* T21 = c.X1b;
        if (!t_ident) {
            op.template mm_stream<false, 0>(t, c.X1b, [&](int i, int j, T acc) { at(c.X2b, i, j) = acc; }, r, r, none, none,
                                            c.TD, c.X1a);
            __syncthreads();
            T21 = c.X2b;
        }
        op.template mm_stream<false, 0>(c.TD, c.X1a, [&](int i, int j, T acc) { at(c.X2a, i, j) = acc; }, r, r, none, none,
                                        T21, c.RB);
        __syncthreads();
        op.template mm_stream<false, 1>(T21, c.RB, [&](int i, int j, T acc) { at(c.X3b, i, j) = acc; }, v1B, v1B,
                                        [&](int i, T s) { J0pn[i] = jp[i] + s; }, none,
                                        t_ident ? nullptr : c.X2a, t_ident ? nullptr : t);
        __syncthreads();
        T* tmp;
        if (t_ident) {
            op.template mv<true>(c.X2a, v1A, [&](int i, T s) { J0m[i] = J0m[i] + s; });
            if (!rpm_zero) {
                for (int e = op.tid; e < n * n; e += CF::NTHR) {
                    const int i = e % n, j = e / n;
                    at(c.X3b, i, j) += Num<T>::flip(rd(r, i, j), dmask[i] ^ dmask[j]);
                }
                __syncthreads();
            }
            tmp = c.TD; c.TD = c.X2a; c.X2a = tmp;
            tmp = c.RB; c.RB = c.X3b; c.X3b = tmp;
        } else {
            op.template mm_stream<true, 1>(c.X2a, t, [&](int i, int j, T acc) { at(c.TD, i, j) = Num<T>::flip(acc, dmask[j]); },
                                           v1A, v1A, [&](int i, T s) { J0m[i] = J0m[i] + s; }, none, c.X3b, t);
            __syncthreads();
            op.template mm_stream<true, 0>(c.X3b, t, [&](int i, int j, T acc) {
                const T rpm = rpm_zero ? 0.0 : Num<T>::flip(rd(r, i, j), dmask[i] ^ dmask[j]);
                at(c.X1b, i, j) = rpm + Num<T>::flip(acc, dmask[j]);
            }, r, r, none, none, nullptr, nullptr);
            __syncthreads();
            tmp = c.RB; c.RB = c.X1b; c.X1b = tmp;
        }
        tmp = c.J0pc; c.J0pc = c.J0pnew; c.J0pnew = tmp;
        return;
    }
    // M1, then its inverse (K = 2: the diagonal only, the announcement of P1 stays valid)
    op.template mm_stream<false, 1>(r, c.RB, [&](int i, int j, T acc) { at(c.X1a, i, j) = acc; ssa = fma(acc, acc, ssa); },
                                    J0pc, J0pc, [&](int i, T s) { v1A[i] = s + jm[i]; }, none, c.RB, r);
    op.finish_inverse(op.neumann_order(op.block_sum(ssa)), c.X1a, c.X2a);
    // P1, then its inverse
    op.template mm_stream<false, 1>(c.RB, r, [&](int i, int j, T acc) { at(c.X1b, i, j) = acc; ssb = fma(acc, acc, ssb); },
                                    jm, jm, [&](int i, T s) { v1B[i] = J0pc[i] + s; }, none, c.TD, c.X1a);
    op.finish_inverse(op.neumann_order(op.block_sum(ssb)), c.X1b, c.X2b);
    // M2: T01_inv = T-- X1a (:226)
    const T* T21 = t_ident ? c.X1b : c.X2b;
    op.template mm_stream<false, 0>(c.TD, c.X1a, [&](int i, int j, T acc) { at(c.X2a, i, j) = acc; }, r, r, none, none,
                                    t_ident ? nullptr : t, t_ident ? nullptr : c.X1b);
    __syncthreads();
    // P2: T21_inv = t X1b (:251)
    if (!t_ident) {
        op.template mm_stream<false, 0>(t, c.X1b, [&](int i, int j, T acc) { at(c.X2b, i, j) = acc; }, r, r, none, none, c.X2a, r);
        __syncthreads();
    }
    // M3: X1a = T01_inv r ; J0- += T01_inv v1A (:229, :237)
    op.template mm_stream<false, 1>(c.X2a, r, [&](int i, int j, T acc) { at(c.X1a, i, j) = acc; }, v1A, v1A,
                                    [&](int i, T s) { J0m[i] = J0m[i] + s; }, none, T21, c.RB);
    __syncthreads();
    // P4: X3b = T21 R+-
    op.template mm_stream<false, 0>(T21, c.RB, [&](int i, int j, T acc) { at(c.X3b, i, j) = acc; }, r, r, none, none, c.X1a, c.TC);
    __syncthreads();
    // M4: R-+ += X1a T++ (:237)
    op.template mm_stream<false, 0>(c.X1a, c.TC, [&](int i, int j, T acc) { at(c.RA, i, j) += acc; }, r, r, none, none,
                                    t_ident ? T21 : c.X3b, t_ident ? c.TC : t);
    __syncthreads();
    // P5: R+-' = r+- + T21 R+- t-- (:265)   [into the free buffer: X2b (t given) / in place X3b (t = I)]
    T* RBn;
    if (t_ident) {
        if (!rpm_zero)
            for (int e = op.tid; e < n * n; e += CF::NTHR) {
                const int i = e % n, j = e / n;
                at(c.X3b, i, j) += Num<T>::flip(rd(r, i, j), dmask[i] ^ dmask[j]);
            }
        RBn = c.X3b;
    } else {
        // X2b (= T21) is still an operand of P3: the new R+- goes to X1b (dead since P2)
        op.template mm_stream<true, 0>(c.X3b, t, [&](int i, int j, T acc) {
            const T rpm = rpm_zero ? 0.0 : Num<T>::flip(rd(r, i, j), dmask[i] ^ dmask[j]);
            at(c.X1b, i, j) = rpm + Num<T>::flip(acc, dmask[j]);
        }, r, r, none, none, T21, c.TC);
        RBn = c.X1b;
        __syncthreads();
    }
    // P3: T++' = T21 T++ ; J0+' = j0+ + T21 v1B (:254, :262)   [into X3b (t given) / X2b (t = I)]
    T* TCn = t_ident ? c.X2b : c.X3b;
    if (t_ident) {
        op.template mm_stream<false, 1>(T21, c.TC, [&](int i, int j, T acc) { at(TCn, i, j) = acc; }, v1B, v1B,
                                        [&](int i, T s) { J0pn[i] = jp[i] + s; }, none, nullptr, nullptr);
    } else {
        op.template mm_stream<false, 1>(T21, c.TC, [&](int i, int j, T acc) { at(TCn, i, j) = acc; }, v1B, v1B,
                                        [&](int i, T s) { J0pn[i] = jp[i] + s; }, none, c.X2a, t);
    }
    __syncthreads();
    // M5: T--' = T01_inv t-- (:240)   [in place; t = I: T--' = X2a]
    if (!t_ident) {
        op.template mm_stream<true, 0>(c.X2a, t, [&](int i, int j, T acc) { at(c.TD, i, j) = Num<T>::flip(acc, dmask[j]); },
                                       r, r, none, none, next_r, next_r ? RBn : nullptr);
        __syncthreads();
    }
    // rotate: composite <- results
    T* tmp;
    if (t_ident) {
        tmp = c.TD; c.TD = c.X2a; c.X2a = tmp;
        tmp = c.TC; c.TC = c.X2b; c.X2b = tmp;       // TCn
        tmp = c.RB; c.RB = c.X3b; c.X3b = tmp;       // RBn
    } else {
        tmp = c.TC; c.TC = c.X3b; c.X3b = tmp;       // TCn
        tmp = c.RB; c.RB = c.X1b; c.X1b = tmp;       // RBn
    }
    tmp = c.J0pc; c.J0pc = c.J0pnew; c.J0pnew = tmp;
}

template <class CF, int MINB>
__global__ void __launch_bounds__(CF::NTHR, MINB) k_adding_stream(ColArgs<double> a, unsigned int total) {
    using T = double;
    constexpr int n = CF::NN, NP = CF::NP, ld = CF::LD;
    StreamSmem<CF> sm;
    T* vb = sm.vb;
    T* J0p = vb;            T* J0m = J0p + NP;  T* J0pn = J0m + NP; T* jps = J0pn + NP;   /* surface / generated j+ */
    T* jms = jps + NP;      T* hv = jms + NP;   T* part = hv + NP;  /* 8 NP */
    T* vec = part + 8 * NP; /* 5 NP */          T* v1A = vec + 5 * NP;  T* v1B = v1A + NP;
    T* jpl = v1B + NP;      T* jml = jpl + NP;  /* added-layer sources staged from the hand-off */
    uint32_t* dmask = reinterpret_cast<uint32_t*>(vb + (size_t)CF::VEC * NP + kMaxComp);
    int* piv = reinterpret_cast<int*>(dmask + NP);
    __shared__ unsigned int sh_item;
    OpsStream<CF> op;
    op.stages = sm.stages; op.vec = vec; op.part = part; op.piv = piv; op.dmask = dmask; op.tid = threadIdx.x; op.bar = 0;
    op.neumann = a.neumann;
    const int tid = threadIdx.x;
    T* slab = a.slab + (size_t)blockIdx.x * a.slab_stride;
    const int nOut = a.nVZA * a.nStokes;
    for (int i = tid; i < NP; i += blockDim.x) dmask[i] = (i < n) ? dmask_of(i, a.nStokes) : 0u;
    unsigned int nextv = 0;
    if (tid == 0) nextv = atomicAdd(a.counter, 1u);
    for (;;) {
        __syncthreads();
        if (tid == 0) sh_item = nextv;
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= total) break;
        if (tid == 0) nextv = atomicAdd(a.counter, 1u);
        const int p = item % a.nS;
        const int mi = item / a.nS;
        const int m = a.m0 + mi;
        StreamComp<CF> c;
        c.RA = slab;               c.RB = slab + CF::MSZ;      c.TC = slab + 2 * CF::MSZ;  c.TD = slab + 3 * CF::MSZ;
        c.X1a = slab + 4 * CF::MSZ; c.X2a = slab + 5 * CF::MSZ; c.X1b = slab + 6 * CF::MSZ; c.X2b = slab + 7 * CF::MSZ;
        c.X3b = slab + 8 * CF::MSZ;
        T* gen_r = slab + 9 * CF::MSZ; T* gen_t = slab + 10 * CF::MSZ;
        c.J0pc = J0p; c.J0pnew = J0pn; c.J0m = J0m; c.v1A = v1A; c.v1B = v1B;
        const bool has_surf = (m == 0) || (a.surf_pad != nullptr);
        const int nsteps = a.Nz + (has_surf ? 1 : 0);
        T jps_sun = 0.0;
        // the added layer of step iz: pointers into the hand-off (scattering), generated (non-scattering), or the surface
        auto layer = [&](int iz, const T*& r, const T*& t, const T*& ljp, const T*& ljm) {
            if (iz == a.Nz) {                         // surface (rpv_surface.jl:51-97 / lambertian_surface.jl:41-95)
                const T ts_end = a.p_tausum[(size_t)a.Nz * a.nS + p];
                const T att = Num<T>::exp_(-ts_end / a.mu0);
                const T rho_l = T(2) * a.albedo;
                if (a.surf_pad) {
                    r = a.surf_pad + (size_t)m * CF::MSZ;
                } else {
                    for (int e = tid; e < n * n; e += blockDim.x) {
                        const int i = e % n, j = e / n;
                        const bool ii = (i % a.nStokes) == 0 && (j % a.nStokes) == 0;
                        gen_r[i + (size_t)j * ld] = (ii ? rho_l : T(0)) * (a.mu[j] * a.wt[j]);
                    }
                    r = gen_r;
                }
                for (int i = tid; i < n; i += blockDim.x) {
                    const T rho = a.surf_rho ? a.surf_rho[(size_t)m * n * n + i + (size_t)a.istart * n]
                                             : (((i % a.nStokes) == 0) ? rho_l : T(0));
                    jps[i] = (i == a.istart) ? att : T(0);
                    jms[i] = a.mu0 * rho * att;
                    if (a.sif0 && !a.surf_rho && m == 0) jms[i] += T(2) * a.sif0[(i % a.nStokes) + (size_t)a.nStokes * (a.F0_off + p)];
                }
                jps_sun = att;
                t = nullptr; ljp = jps; ljm = jms;
            } else if (a.scatter[iz]) {
                const size_t slot = ((size_t)mi * a.n_sched_layers + a.sched_idx[iz]) * a.nS + p;
                r = a.add_r + slot * a.add_mstride;
                t = a.add_t + slot * a.add_mstride;
                for (int i = tid; i < n; i += blockDim.x) { jpl[i] = a.add_jp[slot * n + i]; jml[i] = a.add_jm[slot * n + i]; }
                ljp = jpl; ljm = jml;
            } else {                                  // zero_added_noscat! (rt_helpers.jl:174-180)
                const T tau = a.p_tau[(size_t)iz * a.nS + p];
                for (int e = tid; e < n * n; e += blockDim.x) {
                    const int i = e % n, j = e / n;
                    gen_r[i + (size_t)j * ld] = T(0);
                    gen_t[i + (size_t)j * ld] = (i == j) ? Num<T>::exp_(-tau / a.mu[i]) : T(0);
                }
                for (int i = tid; i < n; i += blockDim.x) { jpl[i] = T(0); jml[i] = T(0); }
                r = gen_r; t = gen_t; ljp = jpl; ljm = jml;
            }
        };
        // the pointer the next step's first product will use as r (for the announcement), when it is known now
        auto next_r_of = [&](int iz) -> const T* {
            if (iz >= nsteps) return nullptr;
            if (iz == a.Nz) return a.surf_pad ? a.surf_pad + (size_t)m * CF::MSZ : nullptr;
            if (!a.scatter[iz] || a.iface[iz] != 3) return nullptr;
            const size_t slot = ((size_t)mi * a.n_sched_layers + a.sched_idx[iz]) * a.nS + p;
            return a.add_r + slot * a.add_mstride;
        };
        // ---- layer 0: copy_added_to_composite! (rt_helpers.jl:204-215)
        {
            const T *r0, *t0p, *ljp, *ljm;
            layer(0, r0, t0p, ljp, ljm);
            __syncthreads();
            const bool sc = a.scatter[0] != 0;
            for (int e = tid; e < n * n; e += blockDim.x) {
                const int i = e % n, j = e / n;
                const size_t o = i + (size_t)j * ld;
                const uint32_t sg = dmask[i] ^ dmask[j];
                const T rv = r0[o], tv = t0p[o];
                c.RA[o] = rv;
                c.TC[o] = tv;
                c.RB[o] = sc ? Num<T>::flip(rv, sg) : T(0);
                c.TD[o] = Num<T>::flip(tv, sg);
            }
            for (int i = tid; i < n; i += blockDim.x) { J0p[i] = ljp[i]; J0m[i] = ljm[i]; }
            __syncthreads();
        }
        for (int iz = 1; iz < nsteps; ++iz) {
            const bool surf = (iz == a.Nz);
            const int lz = surf ? a.Nz - 1 : iz;
            const bool nosrc = surf || a.scatter[lz] == 0;
            const int itf = a.iface[lz];
            const T *r, *t, *ljp, *ljm;
            layer(iz, r, t, ljp, ljm);
            __syncthreads();
            if (itf == 3) {
                add_step11_stream<CF>(op, c, r, t, ljp, ljm, nosrc, surf, next_r_of(iz + 1), a.st_Rmp || a.st_Tpp);
            } else {
                op.drain();
                CompPtrs<T> q{c.RA, c.RB, c.TC, c.TD, const_cast<T*>(r), const_cast<T*>(t), c.X1a, c.X2a, c.X1b};
                interact<T, OpsStream<CF>>(op, itf, q, c.J0pc, c.J0m, ljp, ljm, v1A, nosrc, surf);
                c.RA = q.RA; c.RB = q.RB; c.TC = q.TC; c.TD = q.TD; c.X1a = q.X1; c.X2a = q.X2; c.X1b = q.X3;
            }
            if (surf) {                               // interaction_hdrf!: hdr_J0- = R_surf J0+ + j0-(surf)
                const T* J0pc = c.J0pc;
                op.template mv<false>(r, J0pc, [&](int i, T s) { hv[i] = s + ljm[i]; });
            }
        }
        if (!has_surf) {
            for (int i = tid; i < n; i += blockDim.x) hv[i] = T(0);
            __syncthreads();
        }
        const size_t pbase = ((size_t)m * a.nS + p) * nOut;
        const T* J0pc = c.J0pc;
        export_streams<T>(a, m, p, n, ld, c.RA, c.TC, J0m, J0pc);
        for (int o = tid; o < nOut; o += blockDim.x) {
            const int iv = o % a.nVZA, s = o / a.nVZA;
            const T w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
            const int row = a.vza_row[iv] + s;
            a.part_R[pbase + o] = w * J0m[row];
            a.part_T[pbase + o] = w * J0pc[row];
            a.part_H[pbase + o] = w * hv[row];
        }
        if (m == 0 && a.bhr_uw && tid < a.nStokes) {
            const int s = tid;
            T up = T(0), dw = T(0);
            for (int j = s; j < n; j += a.nStokes) {
                up += hv[j] * a.wt[j] * a.mu[j];
                dw += J0pc[j] * a.wt[j] * a.mu[j];
            }
            a.bhr_uw[s + (size_t)a.nStokes * p] = up;
            a.bhr_dw[s + (size_t)a.nStokes * p] = dw + jps_sun * a.mu[a.istart];
        }
    }
}

// R_surf Diagonal(mu w) per Fourier moment in the padded layout (shared by every spectral point)
template <typename T>
__global__ void k_surf_pad(int n, int ld, int msz, int n_moments, const T* __restrict__ rho, const T* __restrict__ mu,
                           const T* __restrict__ wt, T* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)msz * n_moments) return;
    const int m = (int)(e / msz), o = (int)(e % msz);
    const int i = o % ld, j = o / ld;
    out[e] = (i < n && j < n) ? rho[(size_t)m * n * n + i + (size_t)j * n] * (mu[j] * wt[j]) : T(0);
}

}  // namespace vsm

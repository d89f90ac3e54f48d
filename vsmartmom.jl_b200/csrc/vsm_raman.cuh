// vsm_raman.cuh -- rotational-Raman (RRS) variant of the CoreRT loop (BASELINE config 5), fp64.
//
// Reference: rt_kernel!(::RRS) src/CoreRT/CoreKernel/rt_kernel.jl:352-391 = elemental_inelastic!
// (elemental_inelastic.jl:23-206, 479-603) + elemental! + doubling_helper!(::RRS) (doubling_inelastic.jl:13-164)
// + interaction_helper!(::RRS, ::ScatteringInterface_11) (interaction_inelastic.jl:319-521), driven per Fourier
// moment and layer by rt_run(::RRS) (rt_run.jl:383-517) and closed by postprocessing_vza!(::RRS)
// (postprocessing_vza.jl:121-162).
//
// The inelastic state of the reference is (N,N,nSpec,nRaman): one small matrix chain per (n1, dn) *unit*, where
// the unit at output wavelength n1 reads ELASTIC quantities at its own wavelength n1 and at the donor wavelength
// n0 = n1 + i_shift[dn], never inelastic quantities of another unit.  So the work is split in two kernels per layer:
//
//   k_rrs_elastic   one warp per spectral point: elemental + doubling + interaction of the elastic column, and at
//                   every doubling level / interaction it EXPORTS the operators the units need as zero-padded
//                   NP x NP records (r, t, t gp, gp t, gp r, gp r t, the source vectors; (I-rR)^-1-products ...).
//                   Products that depend on one wavelength only (gp t, gp r t, inv1 r T++ ...) are formed here once
//                   per point instead of once per unit (the reference recomputes them for every dn).
//   k_rrs_inel      one warp per unit: the inelastic elemental layer, nd doubling levels (10 N^3 products + 10
//                   mat-vecs each), the _11 interaction (18 products) against the unit's composite state in HBM.
//
// Every product is an fp64 tensor-core contraction (mma.sync.m8n8k4.f64 = DMMA) issued by ONE warp on operands in
// its private shared-memory arena (ld = NP+4: conflict-free fragment loads) or read as fragments straight from the
// L2-resident records.  No CTA-wide barrier anywhere; __syncwarp only.
#pragma once
#include "vsm_device.cuh"
#include "vsm_raman_args.h"

namespace vsm {
namespace rrs {

constexpr double kWeightTol = 1e-8;   // rt_weight_tol(Float64), rt_helpers.jl:58
constexpr double kCloseTol = 1e-8;    // rt_close_tol, :68
constexpr double kLooseTol = 1e-6;    // rt_loose_tol, :78

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Operand accessors.  A product C = A B runs its k loop in the order k = 16 kb + 4 tg + s (lane tg, step s): any
// order is valid as long as both operands use it, and this one lets a lane fetch its four k-values of a global
// operand as ONE 32-byte run (two 16-byte loads; one fully used sector) instead of four 8-byte gathers:
//   a4(i, k0, v): v[s] = M(i, k0+s)      b4(k0, j, v): v[s] = M(k0+s, j)      operator()(i, j) = M(i, j)
//   SM   shared, column-major, ld = NP+1 (== 1 mod 16: a4 / b4 / mat-vec rows are bank-conflict free)
//   GB   global, column-major ld = NP        -> the B side of a product  (b4 vectorised)
//   GA   global, stored TRANSPOSED (ld = NP) -> the A side of a product  (a4 vectorised)
template <int NP> struct SM {
    const double* p;
    __device__ __forceinline__ double operator()(int i, int j) const { return p[i + j * (NP + 1)]; }
    __device__ __forceinline__ void a4(int i, int k0, double (&v)[4]) const {
#pragma unroll
        for (int s = 0; s < 4; ++s) v[s] = p[i + (k0 + s) * (NP + 1)];
    }
    __device__ __forceinline__ void b4(int k0, int j, double (&v)[4]) const {
#pragma unroll
        for (int s = 0; s < 4; ++s) v[s] = p[k0 + s + j * (NP + 1)];
    }
};
template <int NP, bool RO = true> struct GB {
    const double* p;
    __device__ __forceinline__ double operator()(int i, int j) const { return RO ? __ldg(p + i + j * NP) : p[i + j * NP]; }
    __device__ __forceinline__ void a4(int i, int k0, double (&v)[4]) const {
#pragma unroll
        for (int s = 0; s < 4; ++s) v[s] = (*this)(i, k0 + s);
    }
    __device__ __forceinline__ void b4(int k0, int j, double (&v)[4]) const {
        const double2* q = reinterpret_cast<const double2*>(p + k0 + j * NP);
        const double2 x = RO ? __ldg(q) : q[0], y = RO ? __ldg(q + 1) : q[1];
        v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
    }
};
template <int NP> struct GA {
    const double* p;
    __device__ __forceinline__ double operator()(int i, int j) const { return __ldg(p + j + i * NP); }
    __device__ __forceinline__ void a4(int i, int k0, double (&v)[4]) const {
        const double2* q = reinterpret_cast<const double2*>(p + k0 + i * NP);
        const double2 x = __ldg(q), y = __ldg(q + 1);
        v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
    }
    __device__ __forceinline__ void b4(int k0, int j, double (&v)[4]) const {
#pragma unroll
        for (int s = 0; s < 4; ++s) v[s] = (*this)(k0 + s, j);
    }
};
template <class M> struct Conj {      // D M D
    M m; const double* s;
    __device__ __forceinline__ double operator()(int i, int j) const { return m(i, j) * s[i] * s[j]; }
    __device__ __forceinline__ void a4(int i, int k0, double (&v)[4]) const {
        m.a4(i, k0, v);
        const double si = s[i];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] *= si * s[k0 + q];
    }
    __device__ __forceinline__ void b4(int k0, int j, double (&v)[4]) const {
        m.b4(k0, j, v);
        const double sj = s[j];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] *= sj * s[k0 + q];
    }
};
struct SV { const double* p; __device__ __forceinline__ double operator()(int k) const { return p[k]; } };
struct GV { const double* p; __device__ __forceinline__ double operator()(int k) const { return __ldg(p + k); } };

template <int NP>
__device__ __forceinline__ void czero(double (&c)[NP / 8][NP / 8][2]) {
#pragma unroll
    for (int r = 0; r < NP / 8; ++r)
#pragma unroll
        for (int q = 0; q < NP / 8; ++q) c[r][q][0] = c[r][q][1] = 0.0;
}

// c += A B on the fp64 tensor cores over the zero-padded NP x NP operands (k order: see above)
template <int NP, class FA, class FB>
__device__ __forceinline__ void wmm(double (&c)[NP / 8][NP / 8][2], FA a, FB b, int lane) {
    constexpr int NT = NP / 8;
    const int g = lane >> 2, tg = lane & 3;
    if constexpr (NP <= 16) {
#pragma unroll
        for (int kb = 0; kb < NP; kb += 16) {
            double af[NT][4], bf[NT][4];
#pragma unroll
            for (int r = 0; r < NT; ++r) a.a4(8 * r + g, kb + 4 * tg, af[r]);
#pragma unroll
            for (int q = 0; q < NT; ++q) b.b4(kb + 4 * tg, 8 * q + g, bf[q]);
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int r = 0; r < NT; ++r)
#pragma unroll
                    for (int q = 0; q < NT; ++q) dmma(c[r][q][0], c[r][q][1], af[r][s], bf[q][s]);
        }
    } else {
#pragma unroll 1
        for (int kb = 0; kb < NP; kb += 16)
#pragma unroll 1
            for (int s = 0; s < 4; ++s) {
                double af[NT], bf[NT];
                const int k = kb + 4 * tg + s;
#pragma unroll
                for (int r = 0; r < NT; ++r) af[r] = a(8 * r + g, k);
#pragma unroll
                for (int q = 0; q < NT; ++q) bf[q] = b(k, 8 * q + g);
#pragma unroll
                for (int r = 0; r < NT; ++r)
#pragma unroll
                    for (int q = 0; q < NT; ++q) dmma(c[r][q][0], c[r][q][1], af[r], bf[q]);
            }
    }
}

// f(i, j, c_ij) for the accumulator elements this lane owns
template <int NP, class F>
__device__ __forceinline__ void wepi(double (&c)[NP / 8][NP / 8][2], F f, int lane) {
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int r = 0; r < NP / 8; ++r)
#pragma unroll
        for (int q = 0; q < NP / 8; ++q) {
            f(8 * r + g, 8 * q + 2 * tg, c[r][q][0]);
            f(8 * r + g, 8 * q + 2 * tg + 1, c[r][q][1]);
        }
}

// y[b] = sum_k a(16 b + (lane & 15), k) x(k), complete on every lane (the two half-warps split k)
template <int NP, class FA, class FX>
__device__ __forceinline__ void wmv(double (&y)[NP / 16], FA a, FX x, int lane) {
    const int row = lane & 15, half = lane >> 4;
#pragma unroll
    for (int b = 0; b < NP / 16; ++b) {
        double s = 0.0;
#pragma unroll 2
        for (int k = half * (NP / 2); k < (half + 1) * (NP / 2); k += 4) {
            double v[4];
            a.a4(16 * b + row, k, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) s = fma(v[q], x(k + q), s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        y[b] = s;
    }
}

// I + M (identity on the leading n x n block) as an operand
template <class M> struct IPlus {
    M m; int n;
    __device__ __forceinline__ double operator()(int i, int j) const { return m(i, j) + ((i == j && i < n) ? 1.0 : 0.0); }
    __device__ __forceinline__ void a4(int i, int k0, double (&v)[4]) const {
        m.a4(i, k0, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += (k0 + q == i && i < n) ? 1.0 : 0.0;
    }
    __device__ __forceinline__ void b4(int k0, int j, double (&v)[4]) const {
        m.b4(k0, j, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += (k0 + q == j && j < n) ? 1.0 : 0.0;
    }
};

// ---- register-resident fragments of a 16 x 16 operand (NP = 16): 8 doubles per lane hold the whole matrix as the A side
//      (af[r][s] = M(8r+g, 4tg+s)) or as the B side (bf[q][s] = M(4tg+s, 8q+g)), so an operand that enters several
//      products of a doubling level crosses the L1 data pipe once
typedef double Frag16[2][4];
template <class M> __device__ __forceinline__ void lfA(const M& m, Frag16& f, int g, int tg) { m.a4(g, 4 * tg, f[0]); m.a4(8 + g, 4 * tg, f[1]); }
template <class M> __device__ __forceinline__ void lfB(const M& m, Frag16& f, int g, int tg) { m.b4(4 * tg, g, f[0]); m.b4(4 * tg, 8 + g, f[1]); }
__device__ __forceinline__ void mmf(double (&c)[2][2][2], const Frag16& a, const Frag16& b) {
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 2; ++q) dmma(c[r][q][0], c[r][q][1], a[r][s], b[q][s]);
}
// two mat-vecs that share the matrix as ONE tensor-core tile: B tile column 0 = x1, column 1 = x2, the rest zero;
// on return lanes with tg == 0 hold y1[8r+g] = cv[r][0], y2[8r+g] = cv[r][1]
template <class X1, class X2>
__device__ __forceinline__ void mvf(double (&cv)[2][2], const Frag16& a, X1 x1, X2 x2, int g, int tg) {
    double bx[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) bx[s] = (g == 0) ? x1(4 * tg + s) : ((g == 1) ? x2(4 * tg + s) : 0.0);
    cv[0][0] = cv[0][1] = cv[1][0] = cv[1][1] = 0.0;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int r = 0; r < 2; ++r) dmma(cv[r][0], cv[r][1], a[r][s], bx[s]);
}

// y = M x from the A-side fragments of M on the FMA datapath: 8 DFMA + 4 shuffles per lane (a tensor tile with one useful
// column would spend 8 DMMA = 8x the fp64-pipe time); on return every lane holds y[8r+g] in out[r]
template <class X>
__device__ __forceinline__ void mvs(double (&out)[2], const Frag16& a, X x, int tg) {
    double xs[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) xs[s] = x(4 * tg + s);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        double p = a[r][0] * xs[0];
        p = fma(a[r][1], xs[1], p);
        p = fma(a[r][2], xs[2], p);
        p = fma(a[r][3], xs[3], p);
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        out[r] = p;
    }
}

// in-place inverse of the leading n x n block of M (shared, ld = NP+4) by Gauss-Jordan with partial pivoting
// (the reference's getrf/getri route, cpu_batched.jl:32-37); colk: NP doubles, piv: NP ints of scratch
template <int NP>
__device__ __forceinline__ void winv(double* M, int n, double* colk, int* piv, int lane) {
    constexpr int LDS = NP + 1;
    for (int k = 0; k < n; ++k) {
        __syncwarp();
        double best = -1.0;
        int bi = k;
        for (int i = k + lane; i < n; i += 32) {
            const double v = fabs(M[i + k * LDS]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) piv[k] = bi;
        if (bi != k)
            for (int j = lane; j < n; j += 32) {
                const double t = M[k + j * LDS];
                M[k + j * LDS] = M[bi + j * LDS];
                M[bi + j * LDS] = t;
            }
        __syncwarp();
        const double pinv = 1.0 / M[k + k * LDS];
        for (int i = lane; i < n; i += 32) colk[i] = M[i + k * LDS];
        __syncwarp();
        for (int j = lane; j < n; j += 32) {
            double rk = (j == k) ? pinv : M[k + j * LDS] * pinv;
            M[k + j * LDS] = rk;
            for (int i = 0; i < n; ++i) {
                if (i == k) continue;
                const double base = (j == k) ? 0.0 : M[i + j * LDS];
                M[i + j * LDS] = fma(-colk[i], rk, base);
            }
        }
    }
    __syncwarp();
    for (int k = n - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k)
            for (int i = lane; i < n; i += 32) {
                const double t = M[i + k * LDS];
                M[i + k * LDS] = M[i + p * LDS];
                M[i + p * LDS] = t;
            }
        __syncwarp();
    }
}

// Gm = (I - P)^-1 where the accumulators c hold P.  Adaptive like the large kernels (vsm_device.cuh): with nu = ||P||_F,
// if nu^K / (1 - nu) < 1e-20 the Neumann sum of order K in {2, 4, 8} -- I + P, (I+P)(I+P^2), (I+P)(I+P^2)(I+P^4) --
// equals the inverse to fp64 rounding and runs on the tensor cores; otherwise (or with neumann == 0) Gauss-Jordan with
// partial pivoting.  S and TG are scratch matrices.
template <int NP>
__device__ __forceinline__ void winv_one_minus(double (&c)[NP / 8][NP / 8][2], double* Gm, double* S, double* TG, int n,
                                               int neumann, double* colk, int* piv, int lane) {
    constexpr int LDS = NP + 1;
    double ss = 0.0;
#pragma unroll
    for (int r = 0; r < NP / 8; ++r)
#pragma unroll
        for (int q = 0; q < NP / 8; ++q) ss += c[r][q][0] * c[r][q][0] + c[r][q][1] * c[r][q][1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const double nu = sqrt(ss);
    int K = 0;
    if (neumann && nu < 0.25) {
        const double lim = 1e-20 * (1.0 - nu);
        const double n2 = nu * nu, n4 = n2 * n2;
        K = (n2 < lim) ? 2 : ((n4 < lim) ? 4 : ((n4 * n4 < lim) ? 8 : 0));
    }
    auto idn = [&](int i, int j) { return (i == j && i < n) ? 1.0 : 0.0; };
    if (K == 2) {
        wepi<NP>(c, [&](int i, int j, double v) { Gm[i + j * LDS] = idn(i, j) + v; }, lane);
        __syncwarp();
        return;
    }
    if (K == 0) {
        wepi<NP>(c, [&](int i, int j, double v) { Gm[i + j * LDS] = idn(i, j) - v; }, lane);
        winv<NP>(Gm, n, colk, piv, lane);
        return;
    }
    const SM<NP> sS{S}, sT{TG}, sG{Gm};
    wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);          // S = P
    __syncwarp();
    czero<NP>(c);
    wmm<NP>(c, sS, sS, lane);
    wepi<NP>(c, [&](int i, int j, double v) { TG[i + j * LDS] = v; }, lane);         // TG = P^2
    __syncwarp();
    if (K == 4) {
        czero<NP>(c);
        wmm<NP>(c, IPlus<SM<NP>>{sS, n}, IPlus<SM<NP>>{sT, n}, lane);
        wepi<NP>(c, [&](int i, int j, double v) { Gm[i + j * LDS] = v; }, lane);
        __syncwarp();
        return;
    }
    czero<NP>(c);
    wmm<NP>(c, sT, sT, lane);
    wepi<NP>(c, [&](int i, int j, double v) { Gm[i + j * LDS] = v; }, lane);         // Gm = P^4
    czero<NP>(c);
    wmm<NP>(c, IPlus<SM<NP>>{sS, n}, IPlus<SM<NP>>{sT, n}, lane);
    __syncwarp();
    wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);          // S = (I+P)(I+P^2)
    __syncwarp();
    czero<NP>(c);
    wmm<NP>(c, sS, IPlus<SM<NP>>{sG, n}, lane);
    __syncwarp();
    wepi<NP>(c, [&](int i, int j, double v) { Gm[i + j * LDS] = v; }, lane);
    __syncwarp();
}

// fScattRayleigh = rayl.τ ./ combo.τ before the absorption is added (compEffectiveLayerProperties.jl:56)
static __global__ void k_rrs_fscatt(int nE, int Nz, int nAer, size_t ldspec, const double* __restrict__ tau_rayl,
                             const double* __restrict__ aer_tau_mod, double* __restrict__ fsc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nE) return;
    for (int iz = 0; iz < Nz; ++iz) {
        const double tr = tau_rayl[p + (size_t)iz * ldspec];
        double tot = tr;
        for (int c = 0; c < nAer; ++c) tot = tot + aer_tau_mod[c + nAer * iz];
        fsc[(size_t)iz * nE + p] = tr / tot;
    }
}

// ----------------------------------------------------------------------------------------------------------------
// elastic column with exports: one warp per spectral point.  mode 0: atmospheric layer iz; mode 1: surface step
// ----------------------------------------------------------------------------------------------------------------
template <int NP, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_rrs_elastic(Args a, int m, int iz, int mode) {
    using G = Geo<NP>;
    constexpr int LDS = G::LDS, MS = G::MS, MG = G::MG;
    extern __shared__ __align__(16) unsigned char rrs_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * WPC + warp;
    if (p >= a.nE) return;
    constexpr int kVecs = 12;
    double* base = reinterpret_cast<double*>(rrs_smem) + (size_t)warp * (5 * MS + kVecs * NP);
    double* r = base; double* t = r + MS; double* Gm = t + MS; double* TG = Gm + MS; double* S = TG + MS;
    double* jp = S + MS; double* jm = jp + NP; double* j1p = jm + NP; double* j1m = j1p + NP;
    double* a1 = j1m + NP; double* a2 = a1 + NP; double* colk = a2 + NP; double* wc = colk + NP;   // wc: NP
    int* piv = reinterpret_cast<int*>(wc + NP);                                                     // NP ints (<= NP doubles)
    double* sg = wc + 2 * NP;                                                                       // NP
    const int n = a.N;
    const int nband = a.e0 + p;
    for (int i = lane; i < NP; i += 32) sg[i] = a.sgn[i];
    __syncwarp();
    double c[NP / 8][NP / 8][2];
    const SM<NP> sr{r}, st{t}, sG{Gm}, sTG{TG}, sS{S};
    int nd = 0;

    if (mode == 0) {
        nd = a.ndoubl[iz];
        const size_t pi = (size_t)iz * a.nE + p;
        const double tau = a.p_tau[pi], varpi = a.p_varpi[pi], tausum = a.p_tausum[pi];
        const double dtau = tau / exp2((double)nd);
        double expk = exp(-dtau / a.mu0);
        // ---- elemental! (elemental.jl:289-392, 403-422) with the mixed Z of this point
        const double wsc = (m == 0) ? 0.5 : 0.25;
        const size_t zs = (size_t)n * n;
        const double* Zpp = a.Zpp + (size_t)m * a.nC * zs;
        const double* Zmp = a.Zmp + (size_t)m * a.nC * zs;
        for (int i = lane; i < NP; i += 32) { jp[i] = 0.0; jm[i] = 0.0; }
        for (int e = lane; e < NP * NP; e += 32) {
            const int i = e % NP, j = e / NP;
            double rv = 0.0, tv = 0.0;
            if (i < n && j < n) {
                double zpp = 0.0, zmp = 0.0;
                for (int cc = 0; cc < a.nC; ++cc) {
                    const double w = a.p_W[((size_t)cc * a.Nz + iz) * a.nE + p];
                    zpp = fma(w, Zpp[cc * zs + i + (size_t)j * n], zpp);
                    zmp = fma(w, Zmp[cc * zs + i + (size_t)j * n], zmp);
                }
                const double mui = a.mu[i], muj = a.mu[j], wct = a.wt[j] * wsc;
                if (wct > 2.220446049250313e-16) {
                    rv = varpi * zmp * (muj / (mui + muj)) * wct * (-expm1(-dtau * ((1.0 / mui) + (1.0 / muj))));
                    if (mui == muj) {
                        if (i == j) tv = exp(-dtau / mui) * (1.0 + varpi * zpp * (dtau / mui) * wct);
                        else tv = exp(-dtau / muj) * (varpi * zpp * (dtau / mui) * wct);
                    } else {
                        tv = varpi * zpp * (muj / (mui - muj)) * wct * expdiff_neg<double>(dtau / mui, dtau / muj);
                    }
                } else {
                    tv = (i == j) ? exp(-dtau / mui) : 0.0;
                }
                if (nd >= 1) rv *= sg[i];
            }
            r[i + j * LDS] = rv;
            t[i + j * LDS] = tv;
        }
        {   // get_elem_rt_SFI! (elemental.jl:348-392)
            const double mus = a.mu[a.istart];
            const double att = exp(-tausum / mus);
            for (int i = lane; i < n; i += 32) {
                double zpI = 0.0, zmI = 0.0;
                for (int s = 0; s < a.nStokes; ++s) {
                    const double f0 = a.F0[s + (size_t)a.nStokes * nband];
                    double zpp = 0.0, zmp = 0.0;
                    const int j = a.istart + s;
                    for (int cc = 0; cc < a.nC; ++cc) {
                        const double w = a.p_W[((size_t)cc * a.Nz + iz) * a.nE + p];
                        zpp = fma(w, Zpp[cc * zs + i + (size_t)j * n], zpp);
                        zmp = fma(w, Zmp[cc * zs + i + (size_t)j * n], zmp);
                    }
                    zpI = fma(zpp, f0, zpI);
                    zmI = fma(zmp, f0, zmI);
                }
                const double mui = a.mu[i];
                double vp;
                if (i >= a.istart && i < a.istart + a.nStokes) vp = wsc * varpi * zpI * (dtau / mui) * exp(-dtau / mui);
                else vp = wsc * varpi * zpI * (mus / (mui - mus)) * expdiff_neg<double>(dtau / mui, dtau / mus);
                double vm = wsc * varpi * zmI * (mus / (mui + mus)) * (-expm1(-dtau * ((1.0 / mui) + (1.0 / mus))));
                vp *= att;
                vm *= att;
                if (nd >= 1) vm *= sg[i];
                jp[i] = vp;
                jm[i] = vm;
            }
        }
        __syncwarp();
        // ---- doubling (doubling_inelastic.jl:63-158, elastic lines; rt_helpers.jl:102-166) with the level exports
        for (int l = 0; l < nd; ++l) {
            double* R0 = a.rec0 + ((size_t)l * a.nE + p) * G::REC0;
            double* R1 = a.rec1 + ((size_t)l * a.nE + p) * G::REC1;
            for (int e = lane; e < NP * NP; e += 32) {
                const int i = e % NP, j = e / NP;
                const double rv = r[i + j * LDS];
                R0[e] = rv;
                R0[MG + e] = t[i + j * LDS];
                R1[j + i * NP] = rv;
            }
            czero<NP>(c);
            wmm<NP>(c, sr, sr, lane);
            winv_one_minus<NP>(c, Gm, S, TG, n, a.neumann, colk, piv, lane);
            czero<NP>(c);                                               // tt⁺⁺_gp_refl = t gp
            wmm<NP>(c, st, sG, lane);
            wepi<NP>(c, [&](int i, int j, double v) { TG[i + j * LDS] = v; R1[MG + j + i * NP] = v; }, lane);
            czero<NP>(c);                                               // gp t
            wmm<NP>(c, sG, st, lane);
            wepi<NP>(c, [&](int i, int j, double v) { R0[2 * MG + i + j * NP] = v; }, lane);
            czero<NP>(c);                                               // gp r
            wmm<NP>(c, sG, sr, lane);
            wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; R0[3 * MG + i + j * NP] = v; }, lane);
            __syncwarp();
            czero<NP>(c);                                               // (gp r) t
            wmm<NP>(c, sS, st, lane);
            wepi<NP>(c, [&](int i, int j, double v) { R0[4 * MG + i + j * NP] = v; }, lane);
            // sources (doubling_inelastic.jl:66-73, 111-115)
            for (int i = lane; i < NP; i += 32) { j1p[i] = jp[i] * expk; j1m[i] = jm[i] * expk; }
            __syncwarp();
            double y1[NP / 16], y2[NP / 16];
            wmv<NP>(y1, sr, SV{j1m}, lane);
            wmv<NP>(y2, sr, SV{jp}, lane);
            if (lane < 16)
                for (int b = 0; b < NP / 16; ++b) { a1[16 * b + lane] = jp[16 * b + lane] + y1[b]; a2[16 * b + lane] = j1m[16 * b + lane] + y2[b]; }
            __syncwarp();
            double* V0 = R0 + 5 * MG;
            wmv<NP>(y1, sG, SV{a1}, lane);                              // tmp1 = gp (J0+ + r J1-)
            wmv<NP>(y2, sG, SV{a2}, lane);                              // tmp2 = gp (J1- + r J0+)
            if (lane < 16)
                for (int b = 0; b < NP / 16; ++b) {
                    const int i = 16 * b + lane;
                    V0[i] = y1[b]; V0[NP + i] = y2[b]; V0[2 * NP + i] = jp[i]; V0[3 * NP + i] = j1m[i];
                }
            if (lane == 0) V0[4 * NP] = expk;
            wmv<NP>(y1, sTG, SV{a2}, lane);
            wmv<NP>(y2, sTG, SV{a1}, lane);
            __syncwarp();
            if (lane < 16)
                for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; jm[i] = jm[i] + y1[b]; jp[i] = j1p[i] + y2[b]; }
            // r += (tt r) t ; t = tt t
            czero<NP>(c);
            wmm<NP>(c, sTG, sr, lane);
            wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);
            __syncwarp();
            czero<NP>(c);
            wmm<NP>(c, sS, st, lane);
            wepi<NP>(c, [&](int i, int j, double v) { r[i + j * LDS] += v; }, lane);
            czero<NP>(c);
            wmm<NP>(c, sTG, st, lane);
            __syncwarp();
            wepi<NP>(c, [&](int i, int j, double v) { t[i + j * LDS] = v; }, lane);
            __syncwarp();
            expk = expk * expk;
        }
        // apply_D_matrix! / apply_D_matrix_SFI! (doubling.jl:178-216)
        if (nd >= 1) {
            for (int e = lane; e < NP * NP; e += 32) { const int i = e % NP, j = e / NP; r[i + j * LDS] *= sg[i]; }
            for (int i = lane; i < NP; i += 32) jm[i] *= sg[i];
        }
        __syncwarp();
    } else {
        // ---- create_surface_layer! (lambertian_surface.jl:41-95, rpv_surface.jl:51-97): r-+ = R_surf Diagonal(mu w),
        //      t = I, r+- = 0, j0+ = I0 att, j0- = mu0 R_surf I0 att
        const bool refl = (a.surf_rho != nullptr) && !(a.lambertian && m != 0);
        const double att = exp(-a.p_tausum[(size_t)a.Nz * a.nE + p] / a.mu0);
        const double* rho = refl ? a.surf_rho + (size_t)m * n * n : nullptr;
        for (int e = lane; e < NP * NP; e += 32) {
            const int i = e % NP, j = e / NP;
            const bool in = (i < n && j < n);
            r[i + j * LDS] = (in && refl) ? rho[i + (size_t)j * n] * (a.mu[j] * a.wt[j]) : 0.0;
            t[i + j * LDS] = (in && i == j) ? 1.0 : 0.0;
        }
        for (int i = lane; i < NP; i += 32) {
            const bool src = (i < n) && (refl || !a.lambertian);
            jp[i] = (src && i == a.istart) ? att : 0.0;
            jm[i] = (src && refl) ? a.mu0 * rho[i + (size_t)a.istart * n] * att : 0.0;
        }
        __syncwarp();
    }

    double* Cout = a.cE_out + (size_t)p * G::CE;
    if (mode == 0 && iz == 0) {
        // copy_added_to_composite! (rt_helpers.jl:199-215)
        for (int e = lane; e < NP * NP; e += 32) {
            const int i = e % NP, j = e / NP;
            const double rv = r[i + j * LDS], tv = t[i + j * LDS], s2 = sg[i] * sg[j];
            Cout[MG + e] = rv * s2; Cout[3 * MG + e] = tv * s2;
        }
        for (int i = lane; i < NP; i += 32) { Cout[4 * MG + i] = jp[i]; Cout[4 * MG + NP + i] = jm[i]; }
        return;
    }
    // The composite R-+ / T++ (elastic: interaction_inelastic.jl:489-502, inelastic: ieR-+ / ieT++) enter no J0+- / T-- / R+-
    // line of this or any later step, elastic or inelastic: they are not advanced (slots 0 and 2 MG of the composite and
    // slots 1, 4, 6 MG of the interaction record stay unused), and neither are the products that only feed them.
    // ---- interaction_helper!(::ScatteringInterface_11) elastic lines (interaction_inelastic.jl:372-373, 402-403,
    //      423-427, 441-442, 463-464, 489-502) with the exports the units need
    const double* Cin = a.cE_in + (size_t)p * G::CE;
    const GB<NP, false> Rpm{Cin + MG}, Tmm{Cin + 3 * MG};
    const double* J0p = Cin + 4 * MG; const double* J0m = J0p + NP;
    const bool surf = (mode == 1);
    const Conj<SM<NP>> stmm{st, sg};
    double* I0 = a.irec0 + (size_t)p * G::IREC0;
    double* I1 = a.irec1 + (size_t)p * G::IREC1;
    double* IV = I0 + 8 * MG;
    for (int e = lane; e < NP * NP; e += 32) {
        const int i = e % NP, j = e / NP;
        const double rv = r[i + j * LDS], rpm = Cin[MG + e];
        I0[e] = rpm; I0[2 * MG + e] = rv; I0[3 * MG + e] = t[i + j * LDS] * sg[i] * sg[j];
        I1[2 * MG + j + i * NP] = rv; I1[3 * MG + j + i * NP] = rpm;
    }
    for (int i = lane; i < NP; i += 32) { IV[2 * NP + i] = J0p[i]; IV[3 * NP + i] = jm[i]; }
    double y1[NP / 16], y2[NP / 16];
    // inv1 = (I - r-+ R+-)^-1
    czero<NP>(c);
    wmm<NP>(c, sr, Rpm, lane);
    winv_one_minus<NP>(c, Gm, S, TG, n, a.neumann, colk, piv, lane);
    czero<NP>(c);                                                   // T01_inv = T-- inv1
    wmm<NP>(c, Tmm, sG, lane);
    wepi<NP>(c, [&](int i, int j, double v) { TG[i + j * LDS] = v; I1[j + i * NP] = v; }, lane);
    czero<NP>(c);                                                   // E2 = inv1 t--
    wmm<NP>(c, sG, stmm, lane);
    wepi<NP>(c, [&](int i, int j, double v) { I0[5 * MG + i + j * NP] = v; }, lane);
    wmv<NP>(y1, sr, SV{J0p}, lane);                                 // b1 = j0- + r J0+
    if (lane < 16) for (int b = 0; b < NP / 16; ++b) a1[16 * b + lane] = jm[16 * b + lane] + y1[b];
    __syncwarp();
    wmv<NP>(y1, sG, SV{a1}, lane);                                  // v1 = inv1 b1
    wmv<NP>(y2, sTG, SV{a1}, lane);                                 // J0- += T01 b1
    if (lane < 16)
        for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; IV[i] = y1[b]; Cout[4 * MG + NP + i] = J0m[i] + y2[b]; }
    czero<NP>(c);                                                   // T-- = T01 t--
    wmm<NP>(c, sTG, stmm, lane);
    wepi<NP>(c, [&](int i, int j, double v) { Cout[3 * MG + i + j * NP] = v; }, lane);
    // inv2 = (I - R+- r-+)^-1
    czero<NP>(c);
    wmm<NP>(c, Rpm, sr, lane);
    __syncwarp();
    winv_one_minus<NP>(c, Gm, S, TG, n, a.neumann, colk, piv, lane);
    czero<NP>(c);                                                   // T21_inv = t++ inv2
    wmm<NP>(c, st, sG, lane);
    wepi<NP>(c, [&](int i, int j, double v) { TG[i + j * LDS] = v; I1[MG + j + i * NP] = v; }, lane);
    czero<NP>(c);                                                   // inv2 R+-
    wmm<NP>(c, sG, Rpm, lane);
    wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);
    __syncwarp();
    czero<NP>(c);                                                   // E4 = inv2 R+- t--
    wmm<NP>(c, sS, stmm, lane);
    wepi<NP>(c, [&](int i, int j, double v) { I0[7 * MG + i + j * NP] = v; }, lane);
    wmv<NP>(y1, Rpm, SV{jm}, lane);                                 // b2 = J0+ + R+- j0-
    if (lane < 16) for (int b = 0; b < NP / 16; ++b) a2[16 * b + lane] = J0p[16 * b + lane] + y1[b];
    __syncwarp();
    wmv<NP>(y1, sG, SV{a2}, lane);                                  // v2 = inv2 b2
    wmv<NP>(y2, sTG, SV{a2}, lane);                                 // J0+ = j0+ + T21 b2
    if (lane < 16)
        for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; IV[NP + i] = y1[b]; Cout[4 * MG + i] = jp[i] + y2[b]; }
    czero<NP>(c);                                                   // R+- = r+- + (T21 R+-) t--
    wmm<NP>(c, sTG, Rpm, lane);
    __syncwarp();
    wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);
    __syncwarp();
    czero<NP>(c);
    wmm<NP>(c, sS, stmm, lane);
    wepi<NP>(c, [&](int i, int j, double v) { Cout[MG + i + j * NP] = (surf ? 0.0 : r[i + j * LDS] * sg[i] * sg[j]) + v; }, lane);
    if (surf) {
        // postprocessing_vza!(::RRS) elastic lines (postprocessing_vza.jl:141-142); only owner points are output
        __syncwarp();
        const int q = nband - a.n1lo;
        if (q >= 0 && q < a.n1cnt) {
            const int nOut = a.nVZA * a.nStokes;
            for (int o = lane; o < nOut; o += 32) {
                const int iv = o % a.nVZA, s = o / a.nVZA;
                const double w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
                const int row = a.vza_row[iv] + s;
                a.R_SFI[(size_t)q * nOut + o] += w * Cout[4 * MG + NP + row];
                a.T_SFI[(size_t)q * nOut + o] += w * Cout[4 * MG + row];
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// inelastic units: one warp per (n1, dn)
// ----------------------------------------------------------------------------------------------------------------
template <int NP, int WPC>
__global__ void __launch_bounds__(32 * WPC, (NP <= 16) ? 4 : 1) k_rrs_inel(Args a, int m, int iz, int mode, unsigned int total) {
    using G = Geo<NP>;
    constexpr int LDS = G::LDS, MS = G::MS, MG = G::MG;
    extern __shared__ __align__(16) unsigned char rrs_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * WPC + warp;
    if (item >= total) return;
    int dn, n1;
    if (a.units) { const int2 un = a.units[item]; n1 = un.x; dn = un.y; }      // donor-major, valid units only
    else if (a.order == 0) { dn = item / a.n1cnt; n1 = a.n1lo + item % a.n1cnt; }
    else { dn = item % a.nRaman; n1 = a.u_lo + (int)(item / a.nRaman) - a.shift[dn]; }
    if (n1 < a.n1lo || n1 >= a.n1lo + a.n1cnt) return;
    const int n0 = n1 + a.shift[dn];
    if (n0 < 0 || n0 >= a.nBand) return;                 // get_n₀_n₁ (inelastic_helper.jl:19-26): unit stays zero
    const int p0 = n0 - a.e0, p1 = n1 - a.e0;
    const size_t u = (size_t)dn * a.n1cnt + (n1 - a.n1lo);
    constexpr int kVecs = 7;
    double* base = reinterpret_cast<double*>(rrs_smem) + (size_t)warp * (5 * MS + kVecs * NP);
    double* A = base; double* B = A + MS; double* X = B + MS; double* S1 = X + MS; double* S2 = S1 + MS;
    double* jp = S2 + MS; double* jm = jp + NP; double* v1p = jm + NP; double* v1m = v1p + NP;
    double* w3 = v1m + NP; double* w4 = w3 + NP; double* sg = w4 + NP;
    const int n = a.N;
    for (int i = lane; i < NP; i += 32) sg[i] = a.sgn[i];
    __syncwarp();
    double c[NP / 8][NP / 8][2];
    double y1[NP / 16], y2[NP / 16], y3[NP / 16];
    const SM<NP> sA{A}, sB{B}, sX{X}, sS1{S1}, sS2{S2};

    if (mode == 0) {
        const int nd = a.ndoubl[iz];
        // ---- elemental_inelastic! (elemental_inelastic.jl:122-206, 479-603)
        const double sc = 1.0 / exp2((double)nd);
        const double d1 = a.p_tau[(size_t)iz * a.nE + p1] * sc, d0 = a.p_tau[(size_t)iz * a.nE + p0] * sc;
        const double fw = a.fsc[(size_t)iz * a.nE + p0] * a.wR[dn];
        const double ratio = d1 / d0;
        const double wsc = (m == 0) ? 0.5 : 0.25;
        const double* Zpp = a.ZppR + (size_t)m * n * n;
        const double* Zmp = a.ZmpR + (size_t)m * n * n;
        // Every transcendental of the element formulas depends on one stream at a time: E1_i = expm1(-dτ₁/μ_i),
        // E0_i = expm1(-dτ₀/μ_i), ED_i = exp(-dτ₁/μ_i) - exp(-dτ₀/μ_i) are evaluated once per row (4 N calls instead of
        // 3 N²) and combined exactly: -expm1(-(x+y)) = -(u + v + u v), exp(-y) = 1 + v, and exp(-x) - exp(-y) = u - v
        // unless x and y agree to 1e-3 (then the reference's expdiff_neg itself; cancellation bound 1e3 eps).
        double* vE1 = v1p; double* vE0 = v1m; double* vED = w3; double* va1 = w4; double* va0 = jp;
        for (int i = lane; i < NP; i += 32) {
            const double mui = a.mu[i];
            const double x1 = d1 / mui, x0 = d0 / mui;
            va1[i] = x1; va0[i] = x0;
            vE1[i] = expm1(-x1); vE0[i] = expm1(-x0);
            vED[i] = expdiff_neg<double>(x1, x0);
        }
        __syncwarp();
        auto ediff = [&](int i, int j) {          // exp(-dτ₁/μ_i) - exp(-dτ₀/μ_j)
            const double x = va1[i], y = va0[j];
            if (fabs(x - y) < 1e-3 * fmax(x, y)) return expdiff_neg<double>(x, y);
            return vE1[i] - vE0[j];
        };
        for (int e = lane; e < NP * NP; e += 32) {
            const int i = e % NP, j = e / NP;
            double rv = 0.0, tv = 0.0;
            if (i < n && j < n) {
                const double wct = a.wt[j] * wsc;
                if (wct > kWeightTol) {
                    const double mui = a.mu[i], muj = a.mu[j];
                    const double zpp = Zpp[i + (size_t)j * n], zmp = Zmp[i + (size_t)j * n];
                    const double u = vE1[i], v = vE0[j];
                    const double mr = mui / muj;
                    rv = fw * zmp * (1.0 / (mr + ratio)) * (-(u + v + u * v)) * wct;
                    if (mui == muj) {
                        if (fabs(d0 - d1) > kLooseTol) tv = fw * zpp * wct * vED[i] / (1.0 - ratio);
                        else tv = va0[i] * fw * zpp * wct * (1.0 + v);
                    } else if (fabs(mr - ratio) < kCloseTol) {
                        tv = va0[i] * fw * zpp * wct * (1.0 + v);
                    } else {
                        tv = fw * zpp * (1.0 / (mr - ratio)) * wct * ediff(i, j);
                    }
                    if (nd >= 1) rv *= sg[i];
                }
            }
            A[i + j * LDS] = rv;
            B[i + j * LDS] = tv;
        }
        {
            const double mus = a.mu[a.istart];
            const double att = exp(-a.p_tausum[(size_t)iz * a.nE + p0] / mus);
            const double vs = vE0[a.istart], a0s = va0[a.istart];
            __syncwarp();                // va0 aliases jp: the solar-stream entry is read by every lane before jp is written
            for (int i = lane; i < NP; i += 32) {
                double vp = 0.0, vm = 0.0;
                if (i < n) {
                    double zpI = 0.0, zmI = 0.0;
                    for (int s = 0; s < a.nStokes; ++s) {
                        const double f0 = a.F0[s + (size_t)a.nStokes * n0];
                        zpI = fma(Zpp[i + (size_t)(a.istart + s) * n], f0, zpI);
                        zmI = fma(Zmp[i + (size_t)(a.istart + s) * n], f0, zmI);
                    }
                    const double mui = a.mu[i];
                    const double u = vE1[i];
                    if (i >= a.istart && i < a.istart + a.nStokes) {
                        if (fabs(d0 - d1) > kCloseTol) vp = fw * zpI * wsc * vED[i] / (1.0 - ratio);
                        else vp = va0[i] * wsc * fw * zpI * (1.0 + vE0[i]);
                    } else if (fabs((mui / mus) - ratio) < kCloseTol) {
                        vp = va0[i] * wsc * fw * zpI * (1.0 + vs);
                    } else {
                        const double x = va1[i];
                        const double ed = (fabs(x - a0s) < 1e-3 * fmax(x, a0s)) ? expdiff_neg<double>(x, a0s) : (u - vs);
                        vp = wsc * fw * zpI * (1.0 / ((mui / mus) - ratio)) * ed;
                    }
                    vm = wsc * fw * zmI * (1.0 / ((mui / mus) + ratio)) * (-(u + vs + u * vs));
                    vp *= att;
                    vm *= att;
                    if (nd >= 1) vm *= sg[i];
                }
                jp[i] = vp;
                jm[i] = vm;
            }
        }
        __syncwarp();
        // ---- doubling_helper!(::RRS) inelastic lines (doubling_inelastic.jl:74-106, 118-143)
        for (int l = 0; l < nd; ++l) {
            const double* R0 = a.rec0 + ((size_t)l * a.nE + p0) * G::REC0;
            const double* R1 = a.rec1 + ((size_t)l * a.nE + p1) * G::REC1;
            const GB<NP> r0{R0}, t0{R0 + MG}, GT0{R0 + 2 * MG}, GR0{R0 + 3 * MG}, GRT0{R0 + 4 * MG};
            const GA<NP> r1{R1}, TG1{R1 + MG};
            const double* V0 = R0 + 5 * MG;
            const GV tmp1{V0}, tmp2{V0 + NP}, J0p0{V0 + 2 * NP}, J1m0{V0 + 3 * NP};
            const double e0 = __ldg(V0 + 4 * NP);
            if constexpr (NP == 16) {
                // ---- fragments in registers: every global operand is fetched once per level, the mat-vec pairs ride
                //      on the tensor pipe with the matrix fragments that are already there
                const int g = lane >> 2, tg = lane & 3;
                Frag16 r1A, TG1A, AA, XA, T;
                double cv1[2][2], cv2[2][2], cv3[2][2];
                lfA(r1, r1A, g, tg);
                lfA(TG1, TG1A, g, tg);
                // X = ier r0 + r1 ier
                lfA(sA, AA, g, tg);
                lfB(r0, T, g, tg);
                czero<NP>(c);
                mmf(c, AA, T);
                lfB(sA, T, g, tg);
                mmf(c, r1A, T);
                wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
                if (lane < NP) { v1p[lane] = jp[lane] * e0; v1m[lane] = jm[lane] * e0; }
                __syncwarp();
                lfA(sX, XA, g, tg);
                // sources (:74-98)
#ifdef VSM_RRS_MV_TENSOR
                mvf(cv1, r1A, SV{v1m}, SV{jp}, g, tg);            // r1 ieJ1-          | r1 ieJ0+
                mvf(cv2, AA, J1m0, J0p0, g, tg);                  // ier J1-[n0]       | ier J0+[n0]
                mvf(cv3, XA, tmp1, tmp2, g, tg);                  // X tmp1[n0]        | X tmp2[n0]
#else
                // mat-vecs from the fragments already in registers, on the FMA datapath (mvs)
                { double t0[2], t1[2];
                  mvs(t0, r1A, SV{v1m}, tg); mvs(t1, r1A, SV{jp}, tg);
                  cv1[0][0] = t0[0]; cv1[1][0] = t0[1]; cv1[0][1] = t1[0]; cv1[1][1] = t1[1];
                  mvs(t0, AA, J1m0, tg); mvs(t1, AA, J0p0, tg);
                  cv2[0][0] = t0[0]; cv2[1][0] = t0[1]; cv2[0][1] = t1[0]; cv2[1][1] = t1[1];
                  mvs(t0, XA, tmp1, tg); mvs(t1, XA, tmp2, tg);
                  cv3[0][0] = t0[0]; cv3[1][0] = t0[1]; cv3[0][1] = t1[0]; cv3[1][1] = t1[1]; }
#endif
                if (tg == 0) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int i = 8 * r + g;
                        w3[i] = jp[i] + cv1[r][0] + cv2[r][0] + cv3[r][0];
                        w4[i] = v1m[i] + cv2[r][1] + cv1[r][1] + cv3[r][1];
                    }
                }
                __syncwarp();
                Frag16 BA;
                lfA(sB, BA, g, tg);
#ifdef VSM_RRS_MV_TENSOR
                mvf(cv1, TG1A, SV{w3}, SV{w4}, g, tg);
                mvf(cv2, BA, tmp1, tmp2, g, tg);
#else
                { double t0[2], t1[2];
                  mvs(t0, TG1A, SV{w3}, tg); mvs(t1, TG1A, SV{w4}, tg);
                  cv1[0][0] = t0[0]; cv1[1][0] = t0[1]; cv1[0][1] = t1[0]; cv1[1][1] = t1[1];
                  mvs(t0, BA, tmp1, tg); mvs(t1, BA, tmp2, tg);
                  cv2[0][0] = t0[0]; cv2[1][0] = t0[1]; cv2[0][1] = t1[0]; cv2[1][1] = t1[1]; }
#endif
                if (tg == 0) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int i = 8 * r + g;
                        jp[i] = v1p[i] + cv1[r][0] + cv2[r][0];
                        jm[i] = jm[i] + cv1[r][1] + cv2[r][1];
                    }
                }
                // S1 = iet + X gp t ; S2 = ier + X gp r
                Frag16 GT0B;
                lfB(GT0, GT0B, g, tg);
                czero<NP>(c);
                mmf(c, XA, GT0B);
                wepi<NP>(c, [&](int i, int j, double v) { S1[i + j * LDS] = B[i + j * LDS] + v; }, lane);
                lfB(GR0, T, g, tg);
                czero<NP>(c);
                mmf(c, XA, T);
                wepi<NP>(c, [&](int i, int j, double v) { S2[i + j * LDS] = A[i + j * LDS] + v; }, lane);
                __syncwarp();
                // V = r1 iet + S2 t0   (into X)
                lfB(sB, T, g, tg);
                czero<NP>(c);
                mmf(c, r1A, T);
                lfA(sS2, AA, g, tg);
                lfB(t0, T, g, tg);
                mmf(c, AA, T);
                wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
                __syncwarp();
                // ier' = ier + iet (gp r t)0 + tt1 V
                lfB(GRT0, T, g, tg);
                czero<NP>(c);
                mmf(c, BA, T);
                lfB(sX, T, g, tg);
                mmf(c, TG1A, T);
                wepi<NP>(c, [&](int i, int j, double v) { A[i + j * LDS] += v; }, lane);
                // iet' = tt1 S1 + iet (gp t)0
                lfB(sS1, T, g, tg);
                czero<NP>(c);
                mmf(c, TG1A, T);
                mmf(c, BA, GT0B);
                __syncwarp();
                wepi<NP>(c, [&](int i, int j, double v) { B[i + j * LDS] = v; }, lane);
                __syncwarp();
            } else {
                // X = ier r0 + r1 ier
                czero<NP>(c);
                wmm<NP>(c, sA, r0, lane);
                wmm<NP>(c, r1, sA, lane);
                wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
                for (int i = lane; i < NP; i += 32) { v1p[i] = jp[i] * e0; v1m[i] = jm[i] * e0; }
                __syncwarp();
                // sources (:74-98)
                wmv<NP>(y1, r1, SV{v1m}, lane);
                wmv<NP>(y2, sA, J1m0, lane);
                wmv<NP>(y3, sX, tmp1, lane);
                if (lane < 16) for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; w3[i] = jp[i] + y1[b] + y2[b] + y3[b]; }
                wmv<NP>(y1, sA, J0p0, lane);
                wmv<NP>(y2, r1, SV{jp}, lane);
                wmv<NP>(y3, sX, tmp2, lane);
                if (lane < 16) for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; w4[i] = v1m[i] + y1[b] + y2[b] + y3[b]; }
                __syncwarp();
                wmv<NP>(y1, TG1, SV{w3}, lane);
                wmv<NP>(y2, sB, tmp1, lane);
                wmv<NP>(y3, TG1, SV{w4}, lane);
                double y4[NP / 16];
                wmv<NP>(y4, sB, tmp2, lane);
                __syncwarp();
                if (lane < 16)
                    for (int b = 0; b < NP / 16; ++b) {
                        const int i = 16 * b + lane;
                        jp[i] = v1p[i] + y1[b] + y2[b];
                        jm[i] = jm[i] + y3[b] + y4[b];
                    }
                // S1 = iet + X gp t ; S2 = ier + X gp r
                czero<NP>(c);
                wmm<NP>(c, sX, GT0, lane);
                wepi<NP>(c, [&](int i, int j, double v) { S1[i + j * LDS] = B[i + j * LDS] + v; }, lane);
                czero<NP>(c);
                wmm<NP>(c, sX, GR0, lane);
                wepi<NP>(c, [&](int i, int j, double v) { S2[i + j * LDS] = A[i + j * LDS] + v; }, lane);
                __syncwarp();
                // V = r1 iet + S2 t0   (into X)
                czero<NP>(c);
                wmm<NP>(c, r1, sB, lane);
                wmm<NP>(c, sS2, t0, lane);
                __syncwarp();
                wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
                __syncwarp();
                // ier' = ier + iet (gp r t)0 + tt1 V ;  iet' = tt1 S1 + iet (gp t)0
                czero<NP>(c);
                wmm<NP>(c, sB, GRT0, lane);
                wmm<NP>(c, TG1, sX, lane);
                wepi<NP>(c, [&](int i, int j, double v) { A[i + j * LDS] += v; }, lane);
                czero<NP>(c);
                wmm<NP>(c, TG1, sS1, lane);
                wmm<NP>(c, sB, GT0, lane);
                __syncwarp();
                wepi<NP>(c, [&](int i, int j, double v) { B[i + j * LDS] = v; }, lane);
                __syncwarp();
        }
            }
        // apply_D_IE_RRS! / apply_D_SFI_IE_RRS! (doubling_inelastic.jl:335-358, 396-408)
        if (nd >= 1) {
            for (int e = lane; e < NP * NP; e += 32) { const int i = e % NP, j = e / NP; A[i + j * LDS] *= sg[i]; }
            for (int i = lane; i < NP; i += 32) jm[i] *= sg[i];
        }
        __syncwarp();
    } else {
        // surface step: the surface added layer has no inelastic part
        for (int e = lane; e < MS; e += 32) { A[e] = 0.0; B[e] = 0.0; }
        for (int i = lane; i < NP; i += 32) { jp[i] = 0.0; jm[i] = 0.0; }
        __syncwarp();
    }

    double* Cout = a.cI_out + u * G::CI;
    if (mode == 0 && iz == 0) {
        // copy_added_to_composite_ie! (rt_helpers.jl:222-228)
        for (int e = lane; e < NP * NP; e += 32) {
            const int i = e % NP, j = e / NP;
            const double rv = A[i + j * LDS], tv = B[i + j * LDS], s2 = sg[i] * sg[j];
            Cout[MG + e] = rv * s2; Cout[3 * MG + e] = tv * s2;
        }
        for (int i = lane; i < NP; i += 32) { Cout[4 * MG + i] = jp[i]; Cout[4 * MG + NP + i] = jm[i]; }
        return;
    }
    // (ieR-+ (:400-420) and ieT++ (:461-487) are not advanced, see k_rrs_elastic: nothing reads them back)
    // ---- interaction_helper!(::RRS, ::ScatteringInterface_11) inelastic lines (interaction_inelastic.jl:376-420, 446-487)
    const double* Cin = a.cI_in + u * G::CI;
    const GB<NP> ieTmm{Cin + 3 * MG};
    const double* cJp = Cin + 4 * MG; const double* cJm = cJp + NP;
    const double* I0 = a.irec0 + (size_t)p0 * G::IREC0;
    const double* I1 = a.irec1 + (size_t)p1 * G::IREC1;
    const GB<NP> Rpm0{I0}, r0{I0 + 2 * MG}, tmm0{I0 + 3 * MG}, E2{I0 + 5 * MG}, E4{I0 + 7 * MG};
    const GA<NP> T01{I1}, T21{I1 + MG}, r1{I1 + 2 * MG}, Rpm1{I1 + 3 * MG};
    const double* IV = I0 + 8 * MG;
    const GV v1{IV}, v2{IV + NP}, J0p0{IV + 2 * NP}, aj0m0{IV + 3 * NP};
    const Conj<SM<NP>> sBc{sB, sg};                    // iet-- = D iet++ D
    // the unit's composite ieR+- is used on both sides of products: staged once (coalesced) into S2
    for (int e = lane; e < NP * NP; e += 32) S2[(e % NP) + (e / NP) * LDS] = __ldg(Cin + MG + e);
    const SM<NP> ieRpm{S2};
    __syncwarp();
    if constexpr (NP == 16) {
        // register-resident fragments: T01 / T21, r1 / R+-1, ier, P / P' and ieR+- enter several products each
        const int g = lane >> 2, tg = lane & 3;
        Frag16 TA, RA, AA, PA, T;
        double ya[2], yb[2];
        lfA(T01, TA, g, tg);
        lfA(r1, RA, g, tg);
        lfA(sA, AA, g, tg);
        // Q1 = ier R+-0 + r1 ieR+- ;  P = T01 Q1 + ieT--
        lfB(Rpm0, T, g, tg);
        czero<NP>(c);
        mmf(c, AA, T);
        lfB(ieRpm, T, g, tg);
        mmf(c, RA, T);
        wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
        __syncwarp();
        lfB(sX, T, g, tg);
        czero<NP>(c);
        mmf(c, TA, T);
        wepi<NP>(c, [&](int i, int j, double v) { S1[i + j * LDS] = ieTmm(i, j) + v; }, lane);
        // ieJ0- (:378-398)
        mvs(ya, AA, J0p0, tg);
        mvs(yb, RA, GV{cJp}, tg);
        if (tg == 0) { w3[g] = ya[0] + yb[0] + jm[g]; w3[8 + g] = ya[1] + yb[1] + jm[8 + g]; }
        __syncwarp();
        lfA(sS1, PA, g, tg);
        mvs(ya, TA, SV{w3}, tg);
        mvs(yb, PA, v1, tg);
        if (tg == 0) {
            Cout[4 * MG + NP + g] = __ldg(cJm + g) + ya[0] + yb[0];
            Cout[4 * MG + NP + 8 + g] = __ldg(cJm + 8 + g) + ya[1] + yb[1];
        }
        // ieT-- = T01 iet-- + P E2
        lfB(sBc, T, g, tg);
        czero<NP>(c);
        mmf(c, TA, T);
        lfB(E2, T, g, tg);
        mmf(c, PA, T);
        wepi<NP>(c, [&](int i, int j, double v) { Cout[3 * MG + i + j * NP] = v; }, lane);
        // second half: T21, R+-1, ieR+- (A side)
        lfA(T21, TA, g, tg);
        lfA(Rpm1, RA, g, tg);
        lfA(ieRpm, AA, g, tg);
        // Q3 = ieR+- r0 + R+-1 ier ;  P' = T21 Q3 + iet++
        lfB(r0, T, g, tg);
        czero<NP>(c);
        mmf(c, AA, T);
        lfB(sA, T, g, tg);
        mmf(c, RA, T);
        __syncwarp();
        wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
        __syncwarp();
        lfB(sX, T, g, tg);
        czero<NP>(c);
        mmf(c, TA, T);
        wepi<NP>(c, [&](int i, int j, double v) { S1[i + j * LDS] = B[i + j * LDS] + v; }, lane);
        // ieJ0+ (:447-459)
        mvs(ya, AA, aj0m0, tg);
        mvs(yb, RA, SV{jm}, tg);
        if (tg == 0) { w4[g] = __ldg(cJp + g) + ya[0] + yb[0]; w4[8 + g] = __ldg(cJp + 8 + g) + ya[1] + yb[1]; }
        __syncwarp();
        lfA(sS1, PA, g, tg);
        mvs(ya, TA, SV{w4}, tg);
        mvs(yb, PA, v2, tg);
        if (tg == 0) { Cout[4 * MG + g] = jp[g] + ya[0] + yb[0]; Cout[4 * MG + 8 + g] = jp[8 + g] + ya[1] + yb[1]; }
        // Q4 = ieR+- t--0 + R+-1 iet-- ;  ieR+- = ier+- + T21 Q4 + P' E4
        lfB(tmm0, T, g, tg);
        czero<NP>(c);
        mmf(c, AA, T);
        lfB(sBc, T, g, tg);
        mmf(c, RA, T);
        __syncwarp();
        wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
        __syncwarp();
        lfB(sX, T, g, tg);
        czero<NP>(c);
        mmf(c, TA, T);
        lfB(E4, T, g, tg);
        mmf(c, PA, T);
        wepi<NP>(c, [&](int i, int j, double v) { Cout[MG + i + j * NP] = A[i + j * LDS] * sg[i] * sg[j] + v; }, lane);
    } else {
        // Q1 = ier R+-0 + r1 ieR+- ;  P = T01 Q1 + ieT--
        czero<NP>(c);
        wmm<NP>(c, sA, Rpm0, lane);
        wmm<NP>(c, r1, ieRpm, lane);
        wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
        __syncwarp();
        czero<NP>(c);
        wmm<NP>(c, T01, sX, lane);
        wepi<NP>(c, [&](int i, int j, double v) { S1[i + j * LDS] = ieTmm(i, j) + v; }, lane);
        // ieJ0- (:378-398)
        wmv<NP>(y1, sA, J0p0, lane);
        wmv<NP>(y2, r1, GV{cJp}, lane);
        if (lane < 16) for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; w3[i] = y1[b] + y2[b] + jm[i]; }
        __syncwarp();
        wmv<NP>(y1, T01, SV{w3}, lane);
        wmv<NP>(y2, sS1, v1, lane);
        if (lane < 16) for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; Cout[4 * MG + NP + i] = __ldg(cJm + i) + y1[b] + y2[b]; }
        // ieT-- = T01 iet-- + P E2
        czero<NP>(c);
        wmm<NP>(c, T01, sBc, lane);
        wmm<NP>(c, sS1, E2, lane);
        wepi<NP>(c, [&](int i, int j, double v) { Cout[3 * MG + i + j * NP] = v; }, lane);
        // Q3 = ieR+- r0 + R+-1 ier ;  P' = T21 Q3 + iet++
        czero<NP>(c);
        wmm<NP>(c, ieRpm, r0, lane);
        wmm<NP>(c, Rpm1, sA, lane);
        __syncwarp();
        wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
        __syncwarp();
        czero<NP>(c);
        wmm<NP>(c, T21, sX, lane);
        wepi<NP>(c, [&](int i, int j, double v) { S1[i + j * LDS] = B[i + j * LDS] + v; }, lane);
        // ieJ0+ (:447-459)
        wmv<NP>(y1, ieRpm, aj0m0, lane);
        wmv<NP>(y2, Rpm1, SV{jm}, lane);
        if (lane < 16) for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; w4[i] = __ldg(cJp + i) + y1[b] + y2[b]; }
        __syncwarp();
        wmv<NP>(y1, T21, SV{w4}, lane);
        wmv<NP>(y2, sS1, v2, lane);
        if (lane < 16) for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; Cout[4 * MG + i] = jp[i] + y1[b] + y2[b]; }
        // Q4 = ieR+- t--0 + R+-1 iet-- ;  ieR+- = ier+- + T21 Q4 + P' E4
        czero<NP>(c);
        wmm<NP>(c, ieRpm, tmm0, lane);
        wmm<NP>(c, Rpm1, sBc, lane);
        __syncwarp();
        wepi<NP>(c, [&](int i, int j, double v) { X[i + j * LDS] = v; }, lane);
        __syncwarp();
        czero<NP>(c);
        wmm<NP>(c, T21, sX, lane);
        wmm<NP>(c, sS1, E4, lane);
        wepi<NP>(c, [&](int i, int j, double v) { Cout[MG + i + j * NP] = A[i + j * LDS] * sg[i] * sg[j] + v; }, lane);
    }
    if (mode == 1) {
        // postprocessing_vza!(::RRS) inelastic lines (postprocessing_vza.jl:143-146): this unit's share
        __syncwarp();
        const int nOut = a.nVZA * a.nStokes;
        double* uo = a.uout + u * 2 * nOut;
        for (int o = lane; o < nOut; o += 32) {
            const int iv = o % a.nVZA, s = o / a.nVZA;
            const double w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
            const int row = a.vza_row[iv] + s;
            uo[o] = w * Cout[4 * MG + NP + row];
            uo[nOut + o] = w * Cout[4 * MG + row];
        }
    }
}

// ieR_SFI[i,:,s] += sum_t w ieJ0-[...,s,t] in the reference's order t = 1..nRaman (postprocessing_vza.jl:143-146);
// units whose donor wavelength is outside the band hold zeros in the reference and are skipped here
static __global__ void k_rrs_reduce(Args a, double* __restrict__ ieR, double* __restrict__ ieT, int* __restrict__ nonfinite) {
    const int nOut = a.nVZA * a.nStokes;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)a.n1cnt * nOut) return;
    const int q = (int)(e / nOut), o = (int)(e % nOut);
    double sr = 0.0, st = 0.0;
    for (int dn = 0; dn < a.nRaman; ++dn) {
        const int n0 = a.n1lo + q + a.shift[dn];
        if (n0 < 0 || n0 >= a.nBand) continue;
        const double* uo = a.uout + ((size_t)dn * a.n1cnt + q) * 2 * nOut;
        sr += uo[o];
        st += uo[nOut + o];
    }
    ieR[e] += sr;
    ieT[e] += st;
    if (nonfinite && (!isfinite(ieR[e]) || !isfinite(ieT[e]) || !isfinite(a.R_SFI[e]) || !isfinite(a.T_SFI[e]))) atomicAdd(nonfinite, 1);
}

// ----------------------------------------------------------------------------------------------------------------
// single scattering with rotational Raman: rt_kernel_ss!(::RRS) (CoreKernel/rt_kernel_ss.jl:240-331: elemental_inelastic! of the
// WHOLE layer with ndoubl = 0) + get_interaction_ss_RRS! (CoreKernel/interaction_ss.jl:70-90).  Only source vectors: one warp
// per (output wavelength, moment), lanes over the stream rows, the Raman lines in the reference's order.
// ----------------------------------------------------------------------------------------------------------------

static __global__ void __launch_bounds__(128) k_ss_rrs(SsArgs a, int m, unsigned int total) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int n1 = blockIdx.x * 4 + warp;
    if (n1 >= total) return;
    const int n = a.N;
    const double wsc = (m == 0) ? 0.5 : 0.25;
    const double* Zpp = a.ZppR + (size_t)m * n * n;
    const double* Zmp = a.ZmpR + (size_t)m * n * n;
    const double mus = a.mu[a.istart];
    const int nOut = a.nVZA * a.nStokes;
    for (int i = lane; i < n; i += 32) {
        const double mui = a.mu[i];
        const bool insun = (i >= a.istart && i < a.istart + a.nStokes);
        double sumP = 0.0, sumM = 0.0;
        for (int dn = 0; dn < a.nRaman; ++dn) {
            const int n0 = (int)n1 + a.shift[dn];
            if (n0 < 0 || n0 >= a.nBand) continue;
            double zpI = 0.0, zmI = 0.0;
            for (int s = 0; s < a.nStokes; ++s) {
                const double f0 = a.F0[s + (size_t)a.nStokes * n0];
                zpI = fma(Zpp[i + (size_t)(a.istart + s) * n], f0, zpI);
                zmI = fma(Zmp[i + (size_t)(a.istart + s) * n], f0, zmI);
            }
            double Jp = 0.0, Jm = 0.0;
            for (int iz = 0; iz < a.Nz; ++iz) {
                const double d1 = a.p_tau[(size_t)iz * a.nBand + n1], d0 = a.p_tau[(size_t)iz * a.nBand + n0];
                const double fw = a.fsc[(size_t)iz * a.nBand + n0] * a.wR[dn];
                const double ratio = d1 / d0;
                double vp;
                if (insun) {
                    if (fabs(d0 - d1) > kCloseTol) vp = fw * zpI * wsc * expdiff_neg<double>(d1 / mui, d0 / mui) / (1.0 - ratio);
                    else vp = (d0 / mui) * wsc * fw * zpI * exp(-d0 / mui);
                } else if (fabs((mui / mus) - ratio) < kCloseTol) {
                    vp = (d0 / mui) * wsc * fw * zpI * exp(-d0 / mus);
                } else {
                    vp = wsc * fw * zpI * (1.0 / ((mui / mus) - ratio)) * expdiff_neg<double>(d1 / mui, d0 / mus);
                }
                double vm = wsc * fw * zmI * (1.0 / ((mui / mus) + ratio)) * (-expm1(-((d1 / mui) + (d0 / mus))));
                const double att = exp(-a.p_tausum[(size_t)iz * a.nBand + n0] / mus);
                vp *= att;
                vm *= att;
                if (iz == 0) { Jp = vp; Jm = vm; }
                else {
                    Jp = Jp * exp(-d1 / mui) + vp;
                    Jm = Jm + vm * exp(-a.p_tausum[(size_t)iz * a.nBand + n1] / mui);
                }
            }
            sumP += Jp;
            sumM += Jm;
        }
        for (int iv = 0; iv < a.nVZA; ++iv) {
            const int s = i - a.vza_row[iv];
            if (s >= 0 && s < a.nStokes) {
                const double w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
                a.ieR[(size_t)n1 * nOut + iv + a.nVZA * s] += w * sumM;
                a.ieT[(size_t)n1 * nOut + iv + a.nVZA * s] += w * sumP;
            }
        }
    }
}

template <int NP, int WPC>
constexpr size_t elastic_smem() { return (size_t)WPC * (5 * Geo<NP>::MS + 12 * NP) * sizeof(double) + 64; }
template <int NP, int WPC>
constexpr size_t inel_smem() { return (size_t)WPC * (5 * Geo<NP>::MS + 7 * NP) * sizeof(double) + 64; }

}  // namespace rrs
}  // namespace vsm

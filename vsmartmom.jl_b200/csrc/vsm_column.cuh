// vsm_column.cuh -- fused kernels of the column solver (L2) and layer step (L1).
//
//   k_props   : per-(layer, spectral point) optical-property mixing + tau prefix sum
//               (constructCoreOpticalProperties / extractEffectiveProps,
//                LayerOpticalProperties/compEffectiveLayerProperties.jl:11-93, types.jl:1262-1308)
//   k_layer   : elemental! + ndoubl x doubling! for one (point, layer, moment), all on-chip
//               (CoreKernel/elemental.jl:174-462, doubling.jl:38-252, rt_helpers.jl:102-166)
//   k_adding  : the whole interaction! chain over layers + Lambertian surface + hdrf +
//               postprocessing_vza! for one (point, moment), composite layer on-chip
//               (CoreKernel/interaction.jl:52-266, Surfaces/lambertian_surface.jl:41-95,
//                interaction_hdrf.jl:4-42, tools/postprocessing_vza.jl:23-94)
//   k_reduce  : deterministic sum of per-moment partial outputs (rt_run.jl:383 loop order)
#pragma once
#include "vsm_device.cuh"

namespace vsm {

constexpr int kMaxComp = 8;      // Rayleigh + up to 7 aerosol types

// Debug-only phase timing (-DVSM_PHASE_TIMING): thread 0 of every CTA accumulates clock64()
// deltas per phase of the doubling loop into vsm_phase_cycles[] (read with cudaMemcpyFromSymbol).
#ifdef VSM_PHASE_TIMING
static __device__ unsigned long long vsm_phase_cycles[16];
#define VSM_T0() long long _t0 = clock64()
#define VSM_TICK(k)                                                              \
    do {                                                                         \
        long long _t1 = clock64();                                               \
        if (threadIdx.x == 0) atomicAdd(&vsm_phase_cycles[k], (unsigned long long)(_t1 - _t0)); \
        _t0 = _t1;                                                               \
    } while (0)
static __device__ unsigned long long vsm_add_cycles[16];
#define VSM_A0() long long _a0 = clock64()
#define VSM_ATICK(k)                                                             \
    do {                                                                         \
        long long _a1 = clock64();                                               \
        if (threadIdx.x == 0) atomicAdd(&vsm_add_cycles[k], (unsigned long long)(_a1 - _a0)); \
        _a0 = _a1;                                                               \
    } while (0)
// per-group phases of add2_step11: slots [0..7] group 0, [8..15] group 1
#define VSM_G0() long long _g0 = clock64()
#define VSM_GTICK(k)                                                             \
    do {                                                                         \
        long long _g1 = clock64();                                               \
        if (c.gtid == 0) atomicAdd(&vsm_add_cycles[(k) + 8 * c.group], (unsigned long long)(_g1 - _g0)); \
        _g0 = _g1;                                                               \
    } while (0)
#else
#define VSM_T0()
#define VSM_TICK(k)
#define VSM_A0()
#define VSM_ATICK(k)
#define VSM_G0()
#define VSM_GTICK(k)
#endif

// ----------------------------------------------------------------------------
// Arena: matrix k lives in shared memory if k < n_smem, else in the CTA's slab.
// ----------------------------------------------------------------------------
template <typename T>
struct Arena {
    T* smem;
    T* slab;
    int n_smem;
    size_t msz;
    __device__ T* mat(int k) const {
        return k < n_smem ? smem + (size_t)k * msz : slab + (size_t)(k - n_smem) * msz;
    }
};

// ----------------------------------------------------------------------------
// k_props
// ----------------------------------------------------------------------------
template <typename T>
struct PropsArgs {
    int nS;              // points in this chunk
    int Nz, nAer;
    size_t ldspec;       // stride between layers in tau_rayl / tau_abs (= nSpec of the call)
    const T* tau_rayl;   // already offset to the chunk start
    const T* tau_abs;
    T varpi_cab;
    const T* aer_tau_mod;    // [nAer*Nz] (a + nAer*iz)   createAero τ_mod
    const T* aer_varpi_mod;  // [nAer]
    const int* rayl_allzero; // [Nz]
    // outputs, layout [iz*nS + p]
    T* tau; T* varpi; T* tausum /* (Nz+1) rows */; T* W /* [(c*Nz + iz)*nS + p] */;
};

template <typename T>
__global__ void k_props(PropsArgs<T> a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.nS) return;
    const int nC = a.nAer + 1;
    T ts = T(0);
    a.tausum[p] = ts;
    for (int iz = 0; iz < a.Nz; ++iz) {
        T tau = a.tau_rayl[p + (size_t)iz * a.ldspec];
        T varpi = a.varpi_cab;
        T W[kMaxComp];
        W[0] = T(1);
        bool xz = a.rayl_allzero[iz] != 0;
        for (int c = 0; c < a.nAer; ++c) {
            const T ty = a.aer_tau_mod[c + a.nAer * iz];
            const T py = a.aer_varpi_mod[c];
            const T wx = tau * varpi, wy = ty * py, w = wx + wy;
            const T tnew = tau + ty;
            const T vnew = w / (tnew > T(0) ? tnew : T(1));
            if (xz) {
                for (int q = 0; q <= c; ++q) W[q] = T(0);
                W[c + 1] = T(1);
            } else if (wy == T(0)) {
                W[c + 1] = T(0);
            } else {
                const T fx = wx / w, fy = wy / w;
                for (int q = 0; q <= c; ++q) W[q] *= fx;
                W[c + 1] = fy;
            }
            xz = xz && (wy == T(0));
            tau = tnew;
            varpi = vnew;
        }
        {   // + CoreAbsorptionOpticalProperties (types.jl:1302-1308)
            const T tnew = tau + a.tau_abs[p + (size_t)iz * a.ldspec];
            const T wx = tau * varpi;
            varpi = wx / (tnew > T(0) ? tnew : T(1));
            tau = tnew;
        }
        a.tau[(size_t)iz * a.nS + p] = tau;
        a.varpi[(size_t)iz * a.nS + p] = varpi;
        for (int c = 0; c < nC; ++c) a.W[((size_t)c * a.Nz + iz) * a.nS + p] = W[c];
        ts = ts + tau;                       // τ_sum_all[:,iz+1] = τ_sum_all[:,iz] + τ
        a.tausum[(size_t)(iz + 1) * a.nS + p] = ts;
    }
}

// ----------------------------------------------------------------------------
// shared argument block of k_layer / k_adding
// ----------------------------------------------------------------------------
template <typename T>
struct ColArgs {
    int N, nStokes, Nz, nC, nVZA;
    int ld;                  // arena leading dimension (odd, >= N)
    int istart;              // 0-based first row of the solar stream  (iμ₀Nstart-1)
    int nS;                  // points in this chunk
    int m0, Mb;              // first moment and number of moments of this batch
    int n_sched_layers;      // number of scattering layers (entries of layer_order)
    int n_smem_mats;
    int z_spectral;          // L1 only: Z given per spectral point
    int mz0;                 // moment index of the first Z table (0 for L2; m for L1)
    int neumann;             // 1: adaptive Neumann/Gauss-Jordan inverse (OpsFast fp64), 0: elimination only
    T mu0, albedo;
    const T* surf_rho;       // BRDF surface: Fourier moments of R_surf (N,N,n_moments), or null = Lambertian(albedo)
    const T* mu; const T* wt;                         // [N]
    const int* ndoubl; const int* scatter; const int* iface; const int* layer_order;  // [Nz]
    const int* sched_idx;    // [Nz] position of layer iz in layer_order (scattering layers only)
    const T* Zpp; const T* Zmp;                       // (N,N,nC,n_moments)
    const T* p_tau; const T* p_varpi; const T* p_tausum; const T* p_W;   // k_props outputs
    const T* F0; size_t F0_off;                       // (nStokes, nSpec) ; column offset of chunk
    const T* geo1; const T* geo2;                     // k_geom tables (N,N)
    const T* gs1; const T* gs2;                       // k_geom solar-stream factors [N]
    const unsigned char* gcls;                        // k_geom branch classes (N,N) + [N]
    // added-layer scratch: slot = (mi*n_sched_layers + sched_idx[iz])*nS + p   (scattering layers only)
    T* add_r; T* add_t; T* add_jp; T* add_jm;
    // extra sources (Sources/thermal_emission.jl, Sources/surface_sif.jl), both optional:
    const T* therm_B;        // ThermalEmission.B_layer (Nz, nSpec_total): Planck radiance per layer, volume source at m = 0
    const T* sif0;           // SurfaceSIF.SIF0 (nStokes, nSpec_total): surface emission injected into the surface j0- at m = 0
    size_t add_mstride;      // elements per hand-off matrix: N*N (dense) or LD*NP (streamed kernels, padded layout)
    const T* surf_pad;       // streamed kernels: R_surf Diagonal(mu w) per moment in the padded layout, or null
    // spill slabs and the work counter
    T* slab; size_t slab_stride;
    unsigned int* counter;
    // epilogue
    const int* vza_row;      // [nVZA] 0-based first row of the VZA stream
    const T* vza_w;          // (nStokes,nVZA,n_moments)
    T* part_R; T* part_T; T* part_H;   // [(m*nS + p)*nOut + o]
    T* bhr_uw; T* bhr_dw;              // (nStokes, nS) chunk-local, may be null
    // rt_run_streams export (rt_run.jl:107-193), all optional: the composite after the surface step,
    // per moment, at all streams: (N,N,nSpec_total,n_moments) / (N,nSpec_total,n_moments)
    T* st_Rmp; T* st_Tpp; T* st_Jm; T* st_Jp;
    size_t st_off, st_nspec;           // first spectral index of this chunk, spectral extent of the arrays
};

__device__ inline unsigned int next_item(unsigned int* counter, unsigned int* sh) {
    __syncthreads();
    if (threadIdx.x == 0) *sh = atomicAdd(counter, 1u);
    __syncthreads();
    return *sh;
}

// ----------------------------------------------------------------------------
// elemental r⁻⁺/t⁺⁺ and solar sources into the arena  (elemental.jl:289-392)
// ----------------------------------------------------------------------------
// geometry of the elemental formulas, shared by every point / layer / moment (built once per call):
//   geo1[i,j] = mu_j/(mu_i+mu_j) * w_j ,  geo2[i,j] = mu_j/(mu_i-mu_j) * w_j  (0 where mu_i == mu_j)
//   gcls[i,j] = kind | wflags<<2 : kind 0 distinct streams, 1 nearly coincident (|dmu| <= 1e-3 mu_j, fp32 5e-2: the
//               exact expdiff_neg form is needed), 2 mu_i == mu_j ; wflags bit0/1 = (w_j*0.5 > eps) /
//               (w_j*0.25 > eps), the `wct > eps` test of elemental.jl:296 for m == 0 / m > 0
//   and against the solar stream s = istart:  gs1[i] = mu_s/(mu_i+mu_s), gs2[i] = mu_s/(mu_i-mu_s)
//   (0 inside the solar block), gcls[n*n + i] = 0 distinct / 1 nearly coincident / 2 solar block.
template <typename T>
__global__ void k_geom(int n, int nStokes, int istart, const T* __restrict__ mu, const T* __restrict__ wt,
                       T* __restrict__ geo1, T* __restrict__ geo2, T* __restrict__ gs1, T* __restrict__ gs2,
                       unsigned char* __restrict__ gcls) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * n) return;
    const int i = e % n, j = e / n;
    const T mui = mu[i], muj = mu[j], wj = wt[j];
    const T eps = Num<T>::eps();
    // em1_i - em1_j loses ~mu/|dmu| ulps to cancellation: negligible in fp64 below 1e3 eps, but fp32 keeps the
    // exact expdiff_neg form (the reference's, elemental.jl:319-323) for every pair closer than 5 %
    const T near = (sizeof(T) == 8) ? T(1e-3) : T(5e-2);
    const int kind = (mui == muj) ? 2 : ((Num<T>::abs_(mui - muj) > near * muj) ? 0 : 1);
    geo1[e] = (muj / (mui + muj)) * wj;
    geo2[e] = (kind == 2) ? T(0) : (muj / (mui - muj)) * wj;
    gcls[e] = (unsigned char)(kind | ((wj * T(0.5) > eps) ? 4 : 0) | ((wj * T(0.25) > eps) ? 8 : 0));
    if (e < n) {
        const T mus = mu[istart];
        const bool sun = (e >= istart && e < istart + nStokes);
        const T me = mu[e];
        gs1[e] = mus / (me + mus);
        gs2[e] = sun ? T(0) : mus / (me - mus);
        gcls[n * n + e] = (unsigned char)(sun ? 2 : ((Num<T>::abs_(me - mus) > near * mus) ? 0 : 1));
    }
}

// Scratch: rowtmp >= 3n + 1 (a_i, e_i, em1_i, att) ; cap >= 2*nStokes*n (the solar columns of the mixed,
// scaled Z matrices, captured in the element loop for the source vectors).
template <typename T, int NN, int LDO = 0>
__device__ __forceinline__ void elemental_fill(const ColArgs<T>& a, int m, int nd, T dtau, T varpi,
                                               T tausum, const T* __restrict__ Wc,
                                               const T* __restrict__ Zpp, const T* __restrict__ Zmp,
                                               size_t zstride, const T* __restrict__ F0n,
                                               const uint32_t* __restrict__ dmask,
                                               T* r, T* t, T* jp, T* jm, T* rowtmp, T* cap) {
    const int n = NN ? NN : a.N, ld = LDO ? LDO : (NN ? (NN | 1) : a.ld), nC = a.nC;
    const int nn = n * n;
    const T wsc = (m == 0) ? T(0.5) : T(0.25);       // scaled_weights / fourier_weight
    const unsigned wbit = (m == 0) ? 4u : 8u;
    // per-stream quantities: a_i = dtau/mu_i, e_i = exp(-a_i), em1_i = expm1(-a_i); solar attenuation
    T* ai = rowtmp; T* ei = rowtmp + n; T* em1 = rowtmp + 2 * n; T* satt = rowtmp + 3 * n;
    T* zps = cap; T* zms = cap + a.nStokes * n;
    VSM_T0();
    for (int q = threadIdx.x; q < 2 * n + 1; q += blockDim.x) {
        if (q < n) {
            const T x = dtau / a.mu[q];
            ai[q] = x;
            em1[q] = Num<T>::expm1_(-x);
        } else if (q < 2 * n) {
            ei[q - n] = Num<T>::exp_(-(dtau / a.mu[q - n]));
        } else {
            satt[0] = Num<T>::exp_(-tausum / a.mu[a.istart]);
        }
    }
    __syncthreads();
    VSM_TICK(10);
    // All the fp64 work below competes with the tensor-core products of the co-resident CTAs for
    // the one fp64 datapath, so the element formulas are kept to a minimum of fp64 instructions:
    // the mixing weights carry varpi * w_scale, the geometry and the branch classes are tabulated.
    const T vw = varpi * wsc;
    const T w0 = (nC > 1) ? vw * Wc[0] : vw, w1 = (nC > 1) ? vw * Wc[1] : T(0);   // nC == 1: Z is used as it is
    // element loop, 4 elements per thread per trip: all global (L2) table loads of a trip are issued
    // before any use so their latency overlaps (the arena leaves almost no L1)
    constexpr int U = 4;
    for (int e0 = threadIdx.x; e0 < nn; e0 += U * blockDim.x) {
        T zmv[U], zpv[U], g1v[U], g2v[U];
        unsigned clv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = min(e0 + u * (int)blockDim.x, nn - 1);
            if (nC == 1) {
                zmv[u] = w0 * Zmp[e]; zpv[u] = w0 * Zpp[e];
            } else if (nC == 2) {
                const T a0 = Zmp[e], a1 = Zmp[zstride + e], b0 = Zpp[e], b1 = Zpp[zstride + e];
                zmv[u] = fma(w1, a1, w0 * a0);
                zpv[u] = fma(w1, b1, w0 * b0);
            } else {
                T zm = T(0), zp = T(0);
                for (int c = 0; c < nC; ++c) {
                    zm = fma(Wc[c], Zmp[(size_t)c * zstride + e], zm);
                    zp = fma(Wc[c], Zpp[(size_t)c * zstride + e], zp);
                }
                zmv[u] = vw * zm; zpv[u] = vw * zp;
            }
            g1v[u] = a.geo1[e];
            g2v[u] = a.geo2[e];
            clv[u] = a.gcls[e];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * (int)blockDim.x;
            if (e < nn) {
                const int i = e % n, j = e / n;
                const T zm = zmv[u], zp = zpv[u];          // varpi * w_scale * Z
                const unsigned cl = clv[u];
                if (j >= a.istart && j < a.istart + a.nStokes) {          // solar column: keep for the sources
                    zps[(j - a.istart) * n + i] = zp;
                    zms[(j - a.istart) * n + i] = zm;
                }
                T rv, tv;
                if (cl & wbit) {                                                 // wct > eps (elemental.jl:296)
                    // 1 - exp(-(a_i+a_j)) = -(em1_i + em1_j + em1_i em1_j)      (elemental.jl:298-302)
                    const T xi = em1[i], xj = em1[j];
                    rv = zm * g1v[u] * (-fma(xi, xj, xi + xj));
                    const unsigned kind = cl & 3u;
                    if (kind == 0) {
                        // exp(-a_i) - exp(-a_j) = em1_i - em1_j                    (:319-323)
                        tv = zp * g2v[u] * (xi - xj);
                    } else if (kind == 1) {
                        // nearly coincident streams: the exact expdiff_neg form (rt_helpers.jl:32-40)
                        tv = zp * g2v[u] * expdiff_neg<T>(ai[i], ai[j]);
                    } else {                                                     // mu_i == mu_j (:304-314)
                        const T q = zp * ai[i] * a.wt[j];
                        tv = (i == j) ? ei[i] * (T(1) + q) : ei[j] * q;
                    }
                } else {
                    rv = T(0);
                    tv = (i == j) ? ei[i] : T(0);
                }
                if (nd >= 1) rv = Num<T>::flip(rv, dmask[i]);      // apply_D_elemental!, nd>=1 branch
                r[i + j * ld] = rv;
                t[i + j * ld] = tv;
            }
        }
    }
    __syncthreads();
    VSM_TICK(11);
    // solar sources j0+ / j0- (elemental.jl:343-392) from the captured columns
    const T att = satt[0];
    const T xs = em1[a.istart], as = ai[a.istart];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T zpI = T(0), zmI = T(0);
        for (int s = 0; s < a.nStokes; ++s) {
            const T f = F0n[s];
            zpI = fma(zps[s * n + i], f, zpI);
            zmI = fma(zms[s * n + i], f, zmI);
        }
        const unsigned sc = a.gcls[nn + i];
        const T xi = em1[i];
        T vp;
        if (sc == 2) vp = zpI * ai[i] * ei[i];
        else vp = zpI * a.gs2[i] * ((sc == 0) ? (xi - xs) : expdiff_neg<T>(ai[i], as));
        // -expm1(-dtau (1/mu_i + 1/mu_s)) = -(em1_i + em1_s + em1_i em1_s)
        T vm = zmI * a.gs1[i] * (-fma(xi, xs, xi + xs));
        vp *= att;
        vm *= att;
        if (nd >= 1) vm = Num<T>::flip(vm, dmask[i]);
        jp[i] = vp;
        jm[i] = vm;
    }
    VSM_TICK(12);
}

// ----------------------------------------------------------------------------
// ndoubl doublings in the arena (doubling_helper!, doubling.jl:38-99).
// Four matrices: r, t, X1, X2 (pointers are rotated).
//   X1 = (I - r r)^-1 ; X2 = t X1 ; sources ; r += X2 (r t) ; t = X2 t
// (the reference forms (X2 r) t; the product is re-associated so that one scratch
//  matrix less is live, which lets a third CTA share the SM).
// ----------------------------------------------------------------------------
// sjp / sjm (optional): a per-source slot j0+- with its own expk == 1 (ThermalEmission: the sub-layer's emission is
// not pre-attenuated, doubling.jl:66-80), doubled alongside the solar vectors with the same operators; sv1 / sv2 scratch.
template <typename T, class Ops>
__device__ __forceinline__ void doubling_loop(const Ops& op, int nd, T expk, T*& r, T*& t, T*& X1,
                                              T* X2, T* jp, T* jm, T* j1p, T* j1m, T* v1, T* v2,
                                              T* sjp = nullptr, T* sjm = nullptr, T* sv1 = nullptr, T* sv2 = nullptr) {
    const int n = op.N(), ld = op.LD();
    // the matrix-vector products of the source update ride along with the matrix products that
    // share their left operand (one extra 8-column tile on the tensor-core path)
    auto attenuate = [&]() {                                              // doubling_source_update!, :131-132
        for (int i = op.tid; i < n; i += op.nthreads()) {
            j1p[i] = jp[i] * expk;
            j1m[i] = jm[i] * expk;
        }
    };
    if (nd > 0) {
        attenuate();
        op.sync();
    }
    for (int it = 0; it < nd; ++it) {
        VSM_T0();
        // X1 = (I - r r)^-1 (rt_helpers.jl:103-104) ; v1 = j1m + r jp ; v2 = jp + r j1m  (:133-134)
        op.template inv_one_minus_v<2>(r, r, X1, X2, jp, j1m,
                                       [&](int i, T s) { v1[i] = j1m[i] + s; }, [&](int i, T s) { v2[i] = jp[i] + s; });
        VSM_TICK(2);
        op.template mm<false>(t, X1, [&](int i, int j, T acc) { X2[i + j * ld] = acc; });   // X2 = t gp (:105)
        op.sync();
        if (sjp) {      // doubling_source_update! of the slot (doubling.jl:72-76) with expk = 1: j1 = j
            op.mv2(r, sjp, sjm, [&](int i, T s) { sv1[i] = sjm[i] + s; }, [&](int i, T s) { sv2[i] = sjp[i] + s; });
            op.mv2(X2, sv1, sv2, [&](int i, T s) { sjm[i] = sjm[i] + s; }, [&](int i, T s) { sjp[i] = sjp[i] + s; });
        }
        VSM_TICK(3);
        // doubling_rt_update! (:162-163)
        op.template mm<false>(r, t, [&](int i, int j, T acc) { X1[i + j * ld] = acc; });
        op.sync();
        VSM_TICK(5);
        // r += tt (r t) ; jm += tt v1 ; jp = j1p + tt v2
        op.template mmv<2>(X2, X1, [&](int i, int j, T acc) { r[i + j * ld] += acc; }, v1, v2,
                           [&](int i, T s) { jm[i] = jm[i] + s; }, [&](int i, T s) { jp[i] = j1p[i] + s; });
        op.sync();
        VSM_TICK(6);
        expk = expk * expk;                                               // :164
        op.template mm<false>(X2, t, [&](int i, int j, T acc) { X1[i + j * ld] = acc; });
        if (it + 1 < nd) attenuate();
        op.sync();
        VSM_TICK(7);
        T* tmp = t; t = X1; X1 = tmp;
    }
}

// ThermalEmission volume source of one elemental layer (contribute!, Sources/thermal_emission.jl:236-316):
// j0+-[I rows] = 2 pi (1 - varpi) B_layer[iz, n] (1 - exp(-dtau/mu_i)), zero for Q/U/V rows and zero-mu streams; m = 0 only
template <typename T>
__device__ __forceinline__ void thermal_fill(const ColArgs<T>& a, int n, int iz, int p, T dtau, T varpi, T* sjp, T* sjm) {
    const T B = a.therm_B[iz + (size_t)a.Nz * (a.F0_off + p)];
    const T coeff = T(6.283185307179586) * (T(1) - varpi) * B;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const T mui = a.mu[i];
        const T v = ((i % a.nStokes) == 0 && mui > Num<T>::eps()) ? coeff * (-Num<T>::expm1_(-dtau / mui)) : T(0);
        sjp[i] = v;
        sjm[i] = v;
    }
}

constexpr int kLayerMats = 4;
constexpr int kAddMats = 9;
// vector scratch (in units of n elements of T) shared by both kernels: 6 vectors + vec(5) + part(8) + the
// per-source slot of k_layer (4)
constexpr int kVecUnits = 23;

template <typename T, class Ops, int NN, int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB) k_layer(ColArgs<T> a, unsigned int total) {
    const int n = NN ? NN : a.N;
    const int ld = NN ? (NN | 1) : a.ld;
    const int msz = ld * n;
    T* sm = reinterpret_cast<T*>(vsm_smem);
    const int nsm = NN ? kLayerMats : a.n_smem_mats;
    T* slab = NN ? sm : a.slab + (size_t)blockIdx.x * a.slab_stride;
    auto mat = [&](int k) -> T* { return (NN || k < nsm) ? sm + k * msz : slab + (size_t)(k - nsm) * msz; };
    T* vbase = sm + nsm * msz;
    T* jp = vbase;           T* jm = jp + n;     T* j1p = jm + n;   T* j1m = j1p + n;
    T* v1 = j1m + n;         T* v2 = v1 + n;     T* vec = v2 + n;   /* 5n */
    T* part = vec + 5 * n;   /* 8n */            T* sjp = part + 8 * n;  T* sjm = sjp + n;  T* sv1 = sjm + n;  T* sv2 = sv1 + n;
    T* Wc = sv2 + n;         /* kMaxComp */
    uint32_t* dmask = reinterpret_cast<uint32_t*>(Wc + kMaxComp);
    int* piv = reinterpret_cast<int*>(dmask + n);
    __shared__ unsigned int sh_item;
    Ops op;
    if constexpr (NN == 0) { op.n = n; op.ld = ld; }
    op.vec = vec; op.part = part; op.piv = piv; op.dmask = dmask; op.tid = threadIdx.x; op.bar = 0;
    op.neumann = a.neumann;

    for (int i = threadIdx.x; i < n; i += blockDim.x) dmask[i] = dmask_of(i, a.nStokes);
    unsigned int nextv = 0;
    if (threadIdx.x == 0) nextv = atomicAdd(a.counter, 1u);

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_item = nextv;
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= total) break;
        if (threadIdx.x == 0) nextv = atomicAdd(a.counter, 1u);     // prefetch the next work item
        const int p = item % a.nS;
        const unsigned int rest = item / a.nS;
        const int mi = rest % a.Mb;
        const int si = rest / a.Mb;
        const int iz = a.layer_order[si];
        const int m = a.m0 + mi;
        const int nd = a.ndoubl[iz];
        const size_t pi = (size_t)iz * a.nS + p;
        const T tau = a.p_tau[pi], varpi = a.p_varpi[pi], tausum = a.p_tausum[pi];
        const T dtau = tau / (T)exp2((double)nd);          // dτ = τ ./ 2^ndoubl (rt_kernel.jl:285)
        const T expk = Num<T>::exp_(-dtau / a.mu0);        // init_layer (rt_kernel.jl:347)
        if (threadIdx.x < a.nC) Wc[threadIdx.x] = a.p_W[((size_t)threadIdx.x * a.Nz + iz) * a.nS + p];
        __syncthreads();
        T* r = mat(0); T* t = mat(1); T* X1 = mat(2); T* X2 = mat(3);
        const size_t zstride = (size_t)n * n;
        const T* Zpp = a.Zpp + (size_t)(m - a.mz0) * a.nC * zstride + (a.z_spectral ? p * zstride : 0);
        const T* Zmp = a.Zmp + (size_t)(m - a.mz0) * a.nC * zstride + (a.z_spectral ? p * zstride : 0);
        VSM_T0();
        elemental_fill<T, NN>(a, m, nd, dtau, varpi, tausum, Wc, Zpp, Zmp, zstride,
                              a.F0 + (a.F0_off + p) * a.nStokes, dmask, r, t, jp, jm, part, j1p);
        const bool thermal = (a.therm_B != nullptr) && (m == 0);
        if (thermal) thermal_fill<T>(a, n, iz, p, dtau, varpi, sjp, sjm);
        __syncthreads();
        VSM_TICK(0);
        doubling_loop<T, Ops>(op, nd, expk, r, t, X1, X2, jp, jm, j1p, j1m, v1, v2, thermal ? sjp : nullptr, sjm, sv1, sv2);
        VSM_TICK(8);
        // apply_D_matrix! / apply_D_matrix_SFI! (doubling.jl:178-216): r⁻⁺ <- D r*, j₀⁻ <- D j*
        const size_t slot = ((size_t)mi * a.n_sched_layers + si) * a.nS + p;
        T* gr = a.add_r + slot * zstride;
        T* gt = a.add_t + slot * zstride;
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e % n, j = e / n;
            T rv = r[i + j * ld];
            if (nd >= 1) rv = Num<T>::flip(rv, dmask[i]);
            gr[e] = rv;
            gt[e] = t[i + j * ld];
        }
        // (the interaction step is linear in the sources and treats every slot alike -- interaction.jl:229-259 --
        //  so the doubled per-source slot joins the solar vectors here)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            T vm = jm[i], vp = jp[i];
            if (thermal) { vm += sjm[i]; vp += sjp[i]; }
            if (nd >= 1) vm = Num<T>::flip(vm, dmask[i]);
            a.add_jp[slot * n + i] = vp;
            a.add_jm[slot * n + i] = vm;
        }
        VSM_TICK(9);
    }
}

// rt_run_streams export: composite R-+ / T++ / J0- / J0+ of point p, moment m to global memory
template <typename T>
__device__ __forceinline__ void export_streams(const ColArgs<T>& a, int m, int p, int n, int ld, const T* RA,
                                               const T* TC, const T* J0m, const T* J0p) {
    if (!a.st_Rmp && !a.st_Tpp && !a.st_Jm && !a.st_Jp) return;
    const size_t s = (size_t)m * a.st_nspec + a.st_off + p;
    if (a.st_Rmp || a.st_Tpp) {
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e % n, j = e / n;
            if (a.st_Rmp) a.st_Rmp[s * n * n + e] = RA[i + j * ld];
            if (a.st_Tpp) a.st_Tpp[s * n * n + e] = TC[i + j * ld];
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (a.st_Jm) a.st_Jm[s * n + i] = J0m[i];
        if (a.st_Jp) a.st_Jp[s * n + i] = J0p[i];
    }
}

// ----------------------------------------------------------------------------
// interaction!  (interaction.jl:52-266) on arena-resident composite + added layer.
// Composite: RA=R⁻⁺ RB=R⁺⁻ TC=T⁺⁺ TD=T⁻⁻ ; added: r=r⁻⁺ t=t⁺⁺ (r⁺⁻ = D r D, t⁻⁻ = D t D
// folded in through dmask; rpm_zero / t_ident describe the surface layer where
// r⁺⁻ = 0 and t = I).
// ----------------------------------------------------------------------------
template <typename T>
struct CompPtrs {
    T *RA, *RB, *TC, *TD, *r, *t, *X1, *X2, *X3;
};

template <typename T, class Ops>
__device__ __forceinline__ void interact(const Ops& op, int iface, CompPtrs<T>& c, T* J0p, T* J0m,
                                         const T* jp, const T* jm, T* v1, bool rpm_zero, bool t_ident,
                                         bool full = true) {
    // full == false (ScatteringInterface_11 only): the composite R-+ and T++ are not advanced.  They enter no
    // J0+- / T-- / R+- of this or any later step, so every rt_run output is unchanged; callers that hand the
    // composite matrices out (rt_run_streams, the L1 entry) keep full == true.
    const int n = op.N(), ld = op.LD();
    const uint32_t* dmask = op.dmask;
    auto at = [ld](T* M, int i, int j) -> T& { return M[i + j * ld]; };
    if (iface == 3) {
        // temp2 = I - r⁻⁺ R⁺⁻ ; temp1 = inv ; T01_inv = T⁻⁻ temp1        (:220-226)
        auto build1 = [&]() {
            op.template mm<false>(c.r, c.RB, [&](int i, int j, T acc) {
                at(c.X1, i, j) = (i == j ? T(1) : T(0)) - acc;
            });
        };
        VSM_A0();
        build1();
        __syncthreads();
        VSM_ATICK(1);
        op.inverse(c.X1, build1);
        VSM_ATICK(2);
        op.template mm<false>(c.TD, c.X1, [&](int i, int j, T acc) { at(c.X2, i, j) = acc; });
        // J₀⁻ += T01_inv (r⁻⁺ J₀⁺ + j₀⁻)                                   (:229)
        op.template mv<false>(c.r, J0p, [&](int i, T s) { v1[i] = s + jm[i]; });
        op.template mv<false>(c.X2, v1, [&](int i, T s) { J0m[i] = J0m[i] + s; });
        VSM_ATICK(3);
        // R⁻⁺ += T01_inv r⁻⁺ T⁺⁺                                          (:237)
        if (full) {
            op.template mm<false>(c.X2, c.r, [&](int i, int j, T acc) { at(c.X3, i, j) = acc; });
            __syncthreads();
            op.template mm<false>(c.X3, c.TC, [&](int i, int j, T acc) { at(c.RA, i, j) += acc; });
        }
        // T⁻⁻ = T01_inv t⁻⁻                                               (:240)
        if (t_ident) {
            T* tmp = c.TD; c.TD = c.X2; c.X2 = tmp;
        } else {
            op.template mm<true>(c.X2, c.t, [&](int i, int j, T acc) {
                at(c.TD, i, j) = Num<T>::flip(acc, dmask[j]);
            });
        }
        __syncthreads();
        VSM_ATICK(4);
        // temp2 = I - R⁺⁻ r⁻⁺ ; inv ; T21_inv = t⁺⁺ temp1                  (:248-251)
        auto build2 = [&]() {
            op.template mm<false>(c.RB, c.r, [&](int i, int j, T acc) {
                at(c.X1, i, j) = (i == j ? T(1) : T(0)) - acc;
            });
        };
        build2();
        __syncthreads();
        VSM_ATICK(5);
        op.inverse(c.X1, build2);
        VSM_ATICK(6);
        T* T21;
        if (t_ident) {
            T21 = c.X1;
        } else {
            op.template mm<false>(c.t, c.X1, [&](int i, int j, T acc) { at(c.X2, i, j) = acc; });
            T21 = c.X2;
        }
        // J₀⁺ = j₀⁺ + T21_inv (J₀⁺ + R⁺⁻ j₀⁻)                              (:254)
        op.template mv<false>(c.RB, jm, [&](int i, T s) { v1[i] = J0p[i] + s; });
        op.template mv<false>(T21, v1, [&](int i, T s) { J0p[i] = jp[i] + s; });
        VSM_ATICK(7);
        // T⁺⁺ = T21_inv T⁺⁺                                               (:262)
        if (full) {
            op.template mm<false>(T21, c.TC, [&](int i, int j, T acc) { at(c.X3, i, j) = acc; });
            __syncthreads();
            T* tmp = c.TC; c.TC = c.X3; c.X3 = tmp;
        }
        // R⁺⁻ = r⁺⁻ + T21_inv R⁺⁻ t⁻⁻                                      (:265)
        op.template mm<false>(T21, c.RB, [&](int i, int j, T acc) { at(c.X3, i, j) = acc; });
        __syncthreads();
        if (t_ident) {
            if (rpm_zero) {
                T* tmp = c.RB; c.RB = c.X3; c.X3 = tmp;
            } else {
                for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                    const int i = e % n, j = e / n;
                    at(c.RB, i, j) = Num<T>::flip(at(c.r, i, j), dmask[i] ^ dmask[j]) + at(c.X3, i, j);
                }
            }
        } else {
            op.template mm<true>(c.X3, c.t, [&](int i, int j, T acc) {
                const T rpm = rpm_zero ? T(0) : Num<T>::flip(at(c.r, i, j), dmask[i] ^ dmask[j]);
                at(c.RB, i, j) = rpm + Num<T>::flip(acc, dmask[j]);
            });
        }
        __syncthreads();
        VSM_ATICK(8);
    } else if (iface == 1) {
        // J₀⁻ += T⁻⁻ (r⁻⁺ J₀⁺ + j₀⁻) ; J₀⁺ = j₀⁺ + t⁺⁺ J₀⁺                  (:104-105)
        op.template mv<false>(c.r, J0p, [&](int i, T s) { v1[i] = s + jm[i]; });
        op.template mv<false>(c.TD, v1, [&](int i, T s) { J0m[i] = J0m[i] + s; });
        if (t_ident) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) J0p[i] = jp[i] + J0p[i];
            __syncthreads();
        } else {
            op.template mv<false>(c.t, J0p, [&](int i, T s) { v1[i] = jp[i] + s; });
            for (int i = threadIdx.x; i < n; i += blockDim.x) J0p[i] = v1[i];
            __syncthreads();
        }
        // R⁻⁺ = T⁻⁻ r⁻⁺ T⁺⁺ ; R⁺⁻ = r⁺⁻ ; T⁺⁺ = t⁺⁺ T⁺⁺ ; T⁻⁻ = T⁻⁻ t⁻⁻     (:114-117)
        op.template mm<false>(c.TD, c.r, [&](int i, int j, T acc) { at(c.X3, i, j) = acc; });
        __syncthreads();
        op.template mm<false>(c.X3, c.TC, [&](int i, int j, T acc) { at(c.RA, i, j) = acc; });
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e % n, j = e / n;
            at(c.RB, i, j) = rpm_zero ? T(0) : Num<T>::flip(at(c.r, i, j), dmask[i] ^ dmask[j]);
        }
        if (!t_ident) {
            op.template mm<false>(c.t, c.TC, [&](int i, int j, T acc) { at(c.X1, i, j) = acc; });
            op.template mm<true>(c.TD, c.t, [&](int i, int j, T acc) {
                at(c.X2, i, j) = Num<T>::flip(acc, dmask[j]);
            });
            __syncthreads();
            T* tmp = c.TC; c.TC = c.X1; c.X1 = tmp;
            tmp = c.TD; c.TD = c.X2; c.X2 = tmp;
        }
        __syncthreads();
    } else if (iface == 2) {
        // J₀⁺ = j₀⁺ + t⁺⁺ (J₀⁺ + R⁺⁻ j₀⁻) ; J₀⁻ += T⁻⁻ j₀⁻                  (:149-150)
        op.template mv<false>(c.RB, jm, [&](int i, T s) { v1[i] = J0p[i] + s; });
        if (t_ident) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) J0p[i] = jp[i] + v1[i];
            __syncthreads();
        } else {
            op.template mv<false>(c.t, v1, [&](int i, T s) { J0p[i] = jp[i] + s; });
        }
        op.template mv<false>(c.TD, jm, [&](int i, T s) { J0m[i] = J0m[i] + s; });
        // T⁺⁺ = t⁺⁺ T⁺⁺ ; T⁻⁻ = T⁻⁻ t⁻⁻ ; R⁺⁻ = t⁺⁺ R⁺⁻ t⁻⁻                  (:159-161)
        if (!t_ident) {
            op.template mm<false>(c.t, c.TC, [&](int i, int j, T acc) { at(c.X1, i, j) = acc; });
            op.template mm<true>(c.TD, c.t, [&](int i, int j, T acc) {
                at(c.X2, i, j) = Num<T>::flip(acc, dmask[j]);
            });
            op.template mm<false>(c.t, c.RB, [&](int i, int j, T acc) { at(c.X3, i, j) = acc; });
            __syncthreads();
            T* tmp = c.TC; c.TC = c.X1; c.X1 = tmp;
            tmp = c.TD; c.TD = c.X2; c.X2 = tmp;
            op.template mm<true>(c.X3, c.t, [&](int i, int j, T acc) {
                at(c.RB, i, j) = Num<T>::flip(acc, dmask[j]);
            });
        }
        __syncthreads();
    } else {
        // J₀⁺ = j₀⁺ + t⁺⁺ J₀⁺ ; J₀⁻ += T⁻⁻ j₀⁻ ; T⁻⁻ = t⁻⁻ T⁻⁻ ; T⁺⁺ = t⁺⁺ T⁺⁺    (:60-71)
        if (t_ident) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) J0p[i] = jp[i] + J0p[i];
            __syncthreads();
        } else {
            op.template mv<false>(c.t, J0p, [&](int i, T s) { v1[i] = jp[i] + s; });
            for (int i = threadIdx.x; i < n; i += blockDim.x) J0p[i] = v1[i];
            __syncthreads();
        }
        op.template mv<false>(c.TD, jm, [&](int i, T s) { J0m[i] = J0m[i] + s; });
        if (!t_ident) {
            // t⁻⁻ T⁻⁻ = D t D T⁻⁻ : row sign D_i outside, D_k inside
            op.template mm<true>(c.t, c.TD, [&](int i, int j, T acc) {
                at(c.X2, i, j) = Num<T>::flip(acc, dmask[i]);
            });
            op.template mm<false>(c.t, c.TC, [&](int i, int j, T acc) { at(c.X1, i, j) = acc; });
            __syncthreads();
            T* tmp = c.TC; c.TC = c.X1; c.X1 = tmp;
            tmp = c.TD; c.TD = c.X2; c.X2 = tmp;
        }
        __syncthreads();
    }
}

template <typename T, class Ops, int NN, int NTHR>
__global__ void __launch_bounds__(NTHR) k_adding(ColArgs<T> a, unsigned int total) {
    const int n = NN ? NN : a.N;
    const int ld = NN ? (NN | 1) : a.ld;
    const int msz = ld * n;
    T* sm = reinterpret_cast<T*>(vsm_smem);
    const int nsm = NN ? kAddMats : a.n_smem_mats;
    T* slab = NN ? sm : a.slab + (size_t)blockIdx.x * a.slab_stride;
    auto mat = [&](int k) -> T* { return (NN || k < nsm) ? sm + k * msz : slab + (size_t)(k - nsm) * msz; };
    T* vbase = sm + nsm * msz;
    T* J0p = vbase;         T* J0m = J0p + n;   T* jp = J0m + n;   T* jm = jp + n;
    T* v1 = jm + n;         T* hv = v1 + n;     T* vec = hv + n;   /* 5n */
    T* part = vec + 5 * n;  /* 8n */
    uint32_t* dmask = reinterpret_cast<uint32_t*>(part + 8 * n);
    int* piv = reinterpret_cast<int*>(dmask + n);
    __shared__ unsigned int sh_item;
    const size_t zstride = (size_t)n * n;
    const int nOut = a.nVZA * a.nStokes;
    Ops op;
    if constexpr (NN == 0) { op.n = n; op.ld = ld; }
    op.vec = vec; op.part = part; op.piv = piv; op.dmask = dmask; op.tid = threadIdx.x; op.bar = 0;
    op.neumann = a.neumann;

    for (int i = threadIdx.x; i < n; i += blockDim.x) dmask[i] = dmask_of(i, a.nStokes);
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
        const int mi = item / a.nS;
        const int m = a.m0 + mi;
        CompPtrs<T> c{mat(0), mat(1), mat(2), mat(3), mat(4), mat(5), mat(6), mat(7), mat(8)};
        for (int iz = 0; iz < a.Nz; ++iz) {
            VSM_A0();
            const bool sc = a.scatter[iz] != 0;
            T* dr = (iz == 0) ? c.RA : c.r;
            T* dt = (iz == 0) ? c.TC : c.t;
            T* djp = (iz == 0) ? J0p : jp;
            T* djm = (iz == 0) ? J0m : jm;
            if (sc) {
                const size_t slot = ((size_t)mi * a.n_sched_layers + a.sched_idx[iz]) * a.nS + p;
                const T* gr = a.add_r + slot * zstride;
                const T* gt = a.add_t + slot * zstride;
                for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                    const int i = e % n, j = e / n;
                    dr[i + j * ld] = gr[e];
                    dt[i + j * ld] = gt[e];
                }
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    djp[i] = a.add_jp[slot * n + i];
                    djm[i] = a.add_jm[slot * n + i];
                }
            } else {
                // zero_added_noscat! (rt_helpers.jl:174-180): r = 0, t = diag(exp(-τ/μ)), j = 0
                const T tau = a.p_tau[(size_t)iz * a.nS + p];
                for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                    const int i = e % n, j = e / n;
                    dr[i + j * ld] = T(0);
                    dt[i + j * ld] = (i == j) ? Num<T>::exp_(-tau / a.mu[i]) : T(0);
                }
                for (int i = threadIdx.x; i < n; i += blockDim.x) { djp[i] = T(0); djm[i] = T(0); }
            }
            __syncthreads();
            VSM_ATICK(0);
            if (iz == 0) {
                // copy_added_to_composite! (rt_helpers.jl:204-215)
                for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                    const int i = e % n, j = e / n;
                    const uint32_t s = dmask[i] ^ dmask[j];
                    c.RB[i + j * ld] = sc ? Num<T>::flip(c.RA[i + j * ld], s) : T(0);
                    c.TD[i + j * ld] = Num<T>::flip(c.TC[i + j * ld], s);
                }
                __syncthreads();
            } else {
                interact<T, Ops>(op, a.iface[iz], c, J0p, J0m, jp, jm, v1, !sc, false, a.st_Rmp || a.st_Tpp);
            }
            VSM_ATICK(9);
        }
        // ---- surface (lambertian_surface.jl:41-95 / create_surface_layer!(::AbstractSurfaceType),
        //      rpv_surface.jl:51-97) + last interaction (rt_run.jl:459-465)
        const T ts_end = a.p_tausum[(size_t)a.Nz * a.nS + p];
        T jps_sun = T(0);
        if (m == 0 || a.surf_rho) {
            const T att = Num<T>::exp_(-ts_end / a.mu0);
            const T* rho_m = a.surf_rho ? a.surf_rho + (size_t)m * n * n : nullptr;
            const T rho_l = T(2) * a.albedo;
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                const int i = e % n, j = e / n;
                const bool ii = (i % a.nStokes) == 0 && (j % a.nStokes) == 0;
                const T rho = rho_m ? rho_m[e] : (ii ? rho_l : T(0));
                c.r[i + j * ld] = rho * (a.mu[j] * a.wt[j]);
            }
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const T rho = rho_m ? rho_m[i + (size_t)a.istart * n] : (((i % a.nStokes) == 0) ? rho_l : T(0));
                jp[i] = (i == a.istart) ? att : T(0);
                jm[i] = a.mu0 * rho * att;
                // SurfaceSIF on a Lambertian surface (surface_source_contribute!, Sources/surface_sif.jl:168-178)
                if (a.sif0 && !rho_m && m == 0) jm[i] += T(2) * a.sif0[(i % a.nStokes) + (size_t)a.nStokes * (a.F0_off + p)];
            }
            jps_sun = att;
            __syncthreads();
            interact<T, Ops>(op, a.iface[a.Nz - 1], c, J0p, J0m, jp, jm, v1, true, true, a.st_Rmp || a.st_Tpp);
            // interaction_hdrf!: hdr_J₀⁻ = r⁻⁺ J₀⁺ + j₀⁻
            op.template mv<false>(c.r, J0p, [&](int i, T s) { hv[i] = s + jm[i]; });
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) hv[i] = T(0);
            __syncthreads();
        }
        // ---- rt_run_streams export (streams_callback, rt_run.jl:467-480)
        export_streams<T>(a, m, p, n, ld, c.RA, c.TC, J0m, J0p);
        // ---- postprocessing_vza! into per-moment partials
        const size_t pbase = ((size_t)m * a.nS + p) * nOut;
        for (int o = threadIdx.x; o < nOut; o += blockDim.x) {
            const int iv = o % a.nVZA, s = o / a.nVZA;
            const T w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
            const int row = a.vza_row[iv] + s;
            a.part_R[pbase + o] = w * J0m[row];
            a.part_T[pbase + o] = w * J0p[row];
            a.part_H[pbase + o] = w * hv[row];
        }
        if (m == 0 && a.bhr_uw && threadIdx.x < a.nStokes) {
            const int s = threadIdx.x;
            T up = T(0), dw = T(0);
            for (int j = s; j < n; j += a.nStokes) {
                up += hv[j] * a.wt[j] * a.mu[j];
                dw += J0p[j] * a.wt[j] * a.mu[j];
            }
            a.bhr_uw[s + (size_t)a.nStokes * p] = up;
            a.bhr_dw[s + (size_t)a.nStokes * p] = dw + jps_sun * a.mu[a.istart];
        }
    }
}

// ----------------------------------------------------------------------------
// k_adding2 -- specialised (compile-time N) adding chain.  The two halves of the
// ScatteringInterface_11 step (interaction.jl:220-240 "minus" half: T01, J0-, R-+, T-- ;
// :248-265 "plus" half: T21, J0+, T++, R+-) only read the pre-step composite, so two
// 96-thread groups run them concurrently on named barriers; the plus half writes its
// results to spare buffers and the pointers are rotated at the end of the step.
// Arena: RA RB TC TD | X1a X2a | X1b X2b X3b | r0 t0 r1 t1  (13 matrices, all shared memory).
// The added layer iz+1 is fetched with cp.async into the spare (r,t,j+,j-) set while step iz runs.
// ----------------------------------------------------------------------------
constexpr int kAdd2Mats = 13;
// vector scratch of k_adding2 in units of n: 8 shared vectors + 2 groups x (v1 + vec(5) + part(8))
constexpr int kAdd2VecUnits = 8 + 2 * 14;

template <typename T, int NN, int TM, int GT>
struct Add2Ctx {
    using OpsA = OpsFast<T, TM, NN, GT, 1>;
    using OpsB = OpsFast<T, TM, NN, GT, 2>;
    using OpsC = OpsFast<T, TM, NN, 2 * GT, 0>;
    T *RA, *RB, *TC, *TD, *r, *t, *X1a, *X2a, *X1b, *X2b, *X3b;
    T *J0pc, *J0pnew, *J0m, *jp, *jm, *v1A, *v1B;
    OpsA opA; OpsB opB; OpsC opC;
    const uint32_t* dmask;
    int group, gtid;
};

// one ScatteringInterface_11 step on two groups; ends with the whole CTA synchronised
template <typename T, int NN, int TM, int GT>
__device__ __forceinline__ void add2_step11(Add2Ctx<T, NN, TM, GT>& c, bool rpm_zero, bool t_ident, bool full) {
    constexpr int n = NN, ld = NN | 1;
    const uint32_t* dmask = c.dmask;
    auto at = [](T* M, int i, int j) -> T& { return M[i + j * ld]; };
    VSM_G0();
    if (c.group == 0) {
        const auto& op = c.opA;
        // (re-based on the shared-memory symbol so that the epilogues use LDS/STS, not generic accesses)
        T* r = rebase(c.r); T* RB = rebase(c.RB); T* X1a = rebase(c.X1a); T* X2a = rebase(c.X2a);
        T* TD = rebase(c.TD); T* TC = rebase(c.TC); T* RA = rebase(c.RA); T* t = rebase(c.t);
        T* J0m = rebase(c.J0m); T* v1A = rebase(c.v1A); const T* jm = rebase(c.jm); const T* J0pc = rebase(c.J0pc);
        // (:220-222) with v1A = r J0+ + j0-  (:229) as a by-product
        op.template inv_one_minus_v<1>(r, RB, X1a, X2a, J0pc, J0pc, [&](int i, T s) { v1A[i] = s + jm[i]; }, [](int, T) {});
        VSM_GTICK(0);
        op.template mm<false>(TD, X1a, [&](int i, int j, T acc) { at(X2a, i, j) = acc; });    // T01_inv (:226)
        op.sync();
        VSM_GTICK(1);
        if (full) {
            op.template mmv<1>(X2a, r, [&](int i, int j, T acc) { at(X1a, i, j) = acc; }, v1A, v1A,
                               [&](int i, T s) { J0m[i] = J0m[i] + s; }, [](int, T) {});
            op.sync();
            VSM_GTICK(2);
            op.template mm<false>(X1a, TC, [&](int i, int j, T acc) { at(RA, i, j) += acc; });    // (:237)
        } else {
            // R-+ (and T++ below) of the composite feed no J0+- / T-- / R+- of any later step: without the
            // rt_run_streams export only the source line of (:229) is needed here
            op.template mv<false>(X2a, v1A, [&](int i, T s) { J0m[i] = J0m[i] + s; });
            VSM_GTICK(2);
        }
        if (!t_ident)
            op.template mm<true>(X2a, t, [&](int i, int j, T acc) { at(TD, i, j) = Num<T>::flip(acc, dmask[j]); });  // (:240)
        VSM_GTICK(3);
    } else {
        const auto& op = c.opB;
        T* r = rebase(c.r); T* RB = rebase(c.RB); T* X1b = rebase(c.X1b); T* X2b = rebase(c.X2b);
        T* X3b = rebase(c.X3b); T* TC = rebase(c.TC); T* t = rebase(c.t);
        T* v1B = rebase(c.v1B); T* J0pnew = rebase(c.J0pnew); const T* jm = rebase(c.jm); const T* jp = rebase(c.jp);
        const T* J0pc = rebase(c.J0pc);
        // (:248-249) with v1B = J0+ + R+- j0-  (:254) as a by-product
        op.template inv_one_minus_v<1>(RB, r, X1b, X2b, jm, jm, [&](int i, T s) { v1B[i] = J0pc[i] + s; }, [](int, T) {});
        VSM_GTICK(0);
        T* T21 = X1b;
        if (!t_ident) {
            op.template mm<false>(t, X1b, [&](int i, int j, T acc) { at(X2b, i, j) = acc; });     // T21_inv (:251)
            op.sync();
            T21 = X2b;
        }
        VSM_GTICK(1);
        T* TCn = t_ident ? X2b : X1b;
        if (full) {
            op.template mmv<1>(T21, TC, [&](int i, int j, T acc) { at(TCn, i, j) = acc; }, v1B, v1B,      // (:262)
                               [&](int i, T s) { J0pnew[i] = jp[i] + s; }, [](int, T) {});
            op.template mm<false>(T21, RB, [&](int i, int j, T acc) { at(X3b, i, j) = acc; });
        } else {
            op.template mmv<1>(T21, RB, [&](int i, int j, T acc) { at(X3b, i, j) = acc; }, v1B, v1B,     // (:254)
                               [&](int i, T s) { J0pnew[i] = jp[i] + s; }, [](int, T) {});
        }
        op.sync();
        VSM_GTICK(2);
        if (t_ident) {                                                                // (:265)
            if (!rpm_zero)
                for (int e = c.gtid; e < n * n; e += GT) {
                    const int i = e % n, j = e / n;
                    at(X3b, i, j) += Num<T>::flip(at(r, i, j), dmask[i] ^ dmask[j]);
                }
        } else {
            op.template mm<true>(X3b, t, [&](int i, int j, T acc) {
                const T rpm = rpm_zero ? T(0) : Num<T>::flip(at(r, i, j), dmask[i] ^ dmask[j]);
                at(X2b, i, j) = rpm + Num<T>::flip(acc, dmask[j]);
            });
        }
        VSM_GTICK(3);
    }
    __syncthreads();
    VSM_GTICK(4);
    T* tmp;
    if (t_ident) {
        tmp = c.TD; c.TD = c.X2a; c.X2a = tmp;
        if (full) { tmp = c.TC; c.TC = c.X2b; c.X2b = tmp; }
        tmp = c.RB; c.RB = c.X3b; c.X3b = tmp;
    } else {
        if (full) { tmp = c.TC; c.TC = c.X1b; c.X1b = tmp; }
        tmp = c.RB; c.RB = c.X2b; c.X2b = tmp;
    }
    tmp = c.J0pc; c.J0pc = c.J0pnew; c.J0pnew = tmp;
}

// the rare interfaces (non-scattering layers) run on the whole CTA
template <typename T, int NN, int TM, int GT>
__device__ __forceinline__ void add2_step_other(Add2Ctx<T, NN, TM, GT>& c, int iface, bool rpm_zero, bool t_ident) {
    CompPtrs<T> q{c.RA, c.RB, c.TC, c.TD, c.r, c.t, c.X1a, c.X2a, c.X1b};
    interact<T, typename Add2Ctx<T, NN, TM, GT>::OpsC>(c.opC, iface, q, c.J0pc, c.J0m, c.jp, c.jm, c.v1A, rpm_zero, t_ident);
    c.RA = q.RA; c.RB = q.RB; c.TC = q.TC; c.TD = q.TD; c.X1a = q.X1; c.X2a = q.X2; c.X1b = q.X3;
}

template <typename T, int NN, int TM, int GT>
__global__ void __launch_bounds__(2 * GT, 1) k_adding2(ColArgs<T> a, unsigned int total) {
    constexpr int n = NN, ld = NN | 1, msz = ld * n;
    constexpr int EPV = 16 / (int)sizeof(T);          // elements per 16-byte cp.async
    constexpr int LB = msz + EPV - 1;                 // layer-buffer stride: room to match the source's 16-B phase
    T* sm = reinterpret_cast<T*>(vsm_smem);
    T* vb = sm + 9 * msz + 4 * LB;
    T* J0p = vb;            T* J0m = J0p + n;   T* J0pn = J0m + n;  T* jpb = J0pn + n;   // jpb: 2 x n
    T* jmb = jpb + 2 * n;   T* hv = jmb + 2 * n;                                          // jmb: 2 x n
    T* v1A = hv + n;        T* vecA = v1A + n;  T* partA = vecA + 5 * n;
    T* v1B = partA + 8 * n; T* vecB = v1B + n;  T* partB = vecB + 5 * n;
    uint32_t* dmask = reinterpret_cast<uint32_t*>(partB + 8 * n);
    int* pivA = reinterpret_cast<int*>(dmask + n);
    int* pivB = pivA + n + 4;
    int* lflag = pivB + n + 4;        // [Nz] scatter | iface << 1, staged once (global reads would sit on the step's critical path)
    __shared__ unsigned int sh_item;
    const int tid = threadIdx.x;
    Add2Ctx<T, NN, TM, GT> c;
    c.group = tid / GT; c.gtid = tid % GT; c.dmask = dmask;
    c.opA.vec = vecA; c.opA.part = partA; c.opA.piv = pivA; c.opA.dmask = dmask; c.opA.tid = c.gtid; c.opA.bar = 1;
    c.opB.vec = vecB; c.opB.part = partB; c.opB.piv = pivB; c.opB.dmask = dmask; c.opB.tid = c.gtid; c.opB.bar = 2;
    c.opC.vec = vecA; c.opC.part = partA; c.opC.piv = pivA; c.opC.dmask = dmask; c.opC.tid = tid;   c.opC.bar = 0;
    c.opA.neumann = c.opB.neumann = c.opC.neumann = a.neumann;
    c.J0m = J0m; c.v1A = v1A; c.v1B = v1B;
    const size_t zstride = (size_t)n * n;
    const int nOut = a.nVZA * a.nStokes;
    T* const lbuf = sm + 9 * msz;   // r0 t0 r1 t1, stride LB
    static_assert(kAdd2Mats == 13, "arena: 9 matrices + 4 layer buffers");

    // One matrix of the k_layer hand-off -> shared memory as one TMA bulk copy.  The slots are only
    // element-aligned ((N*N)-strided), so the destination is shifted by up to EPV-1 elements to the
    // source's 16-byte phase; the few head/tail elements go by element-wise cp.async.  Thread 0 issues
    // the bulk copy on `bar` (expect_tx already posted).  Returns the shifted destination.
    __shared__ __align__(8) unsigned long long mbar[2];
    auto matrix_bulk_bytes = [&](const T* g) -> unsigned {
        const int ps = (int)((reinterpret_cast<uintptr_t>(g) / sizeof(T)) % EPV);
        return (unsigned)(((n * n - (EPV - ps) % EPV) / EPV) * 16);
    };
    auto copy_matrix = [&](T* dbase, const T* g, unsigned long long* bar) -> T* {
        const int ps = (int)((reinterpret_cast<uintptr_t>(g) / sizeof(T)) % EPV);
        const int pd = (int)((dbase - sm) % EPV);
        T* d = dbase + ((ps - pd + EPV) % EPV);
        const int head = (EPV - ps) % EPV;
        const int nvec = (n * n - head) / EPV;
        const int tail0 = head + nvec * EPV;
        if (tid == 0) bulk_g2s(d + head, g + head, (unsigned)nvec * 16u, bar);
        if (tid >= 32 && tid - 32 < head) cp_async_elem(d + tid - 32, g + tid - 32);
        if (tid >= 64 && tid - 64 < n * n - tail0) cp_async_elem(d + tail0 + tid - 64, g + tail0 + tid - 64);
        return d;
    };
    // fill (r, t, j+, j-) of layer iz into layer buffer b (or, for iz == 0, into the composite):
    // scattering layers asynchronously from the k_layer hand-off, the others in place.  The
    // destination must not be read by anyone until the next cp_async_wait + __syncthreads.
    T* lr0 = lbuf; T* lt0 = lbuf + LB; T* lr1 = lbuf + 2 * LB; T* lt1 = lbuf + 3 * LB;
    bool pend0 = false, pend1 = false;       // a bulk copy is in flight on mbar[b]   (CTA-uniform)
    unsigned par0 = 0, par1 = 0;             // phase parity the next wait on mbar[b] expects
    // everything fetched into buffer b has landed and is visible to the whole CTA
    auto layer_ready = [&](int b) {
        cp_async_wait_all();
        if (b ? pend1 : pend0) {
            mbar_wait(&mbar[b], b ? par1 : par0);
            if (b) { par1 ^= 1u; pend1 = false; } else { par0 ^= 1u; pend0 = false; }
        }
        __syncthreads();
    };
    auto fetch_layer = [&](int iz, int p, int mi, int b, T* dr0, T* dt0, T* djp, T* djm) {
        T* dr = dr0 ? dr0 : lbuf + (2 * b) * LB;
        T* dt = dt0 ? dt0 : lbuf + (2 * b + 1) * LB;
        if (lflag[iz] & 1) {
            const size_t slot = ((size_t)mi * a.n_sched_layers + a.sched_idx[iz]) * a.nS + p;
            const T* gr = a.add_r + slot * zstride;
            const T* gt = a.add_t + slot * zstride;
            if (dr0 || ld != n) {                        // fixed destination / padded layout: element copies
                for (int e = tid; e < n * n; e += 2 * GT) {
                    const int i = e % n, j = e / n;
                    cp_async_elem(dr + i + j * ld, gr + e);
                    cp_async_elem(dt + i + j * ld, gt + e);
                }
            } else {
                // ld == n: the buffer is a dense copy of the slot.  The buffer was last touched through the
                // generic proxy (reads of the step before last, ordered by the CTA barrier): fence, then TMA.
                if (tid == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&mbar[b], matrix_bulk_bytes(gr) + matrix_bulk_bytes(gt));
                }
                dr = copy_matrix(dr, gr, &mbar[b]);
                dt = copy_matrix(dt, gt, &mbar[b]);
                if (b) pend1 = true; else pend0 = true;
            }
            for (int i = tid; i < n; i += 2 * GT) {
                cp_async_elem(djp + i, a.add_jp + slot * n + i);
                cp_async_elem(djm + i, a.add_jm + slot * n + i);
            }
        } else {
            const T tau = a.p_tau[(size_t)iz * a.nS + p];
            for (int e = tid; e < n * n; e += 2 * GT) {
                const int i = e % n, j = e / n;
                dr[i + j * ld] = T(0);
                dt[i + j * ld] = (i == j) ? Num<T>::exp_(-tau / a.mu[i]) : T(0);
            }
            for (int i = tid; i < n; i += 2 * GT) { djp[i] = T(0); djm[i] = T(0); }
        }
        if (!dr0) {
            if (b) { lr1 = dr; lt1 = dt; } else { lr0 = dr; lt0 = dt; }
        }
    };

    for (int i = tid; i < n; i += blockDim.x) dmask[i] = dmask_of(i, a.nStokes);
    for (int i = tid; i < a.Nz; i += blockDim.x) lflag[i] = (a.scatter[i] != 0 ? 1 : 0) | (a.iface[i] << 1);
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        c.RA = sm;             c.RB = sm + msz;      c.TC = sm + 2 * msz;  c.TD = sm + 3 * msz;
        c.X1a = sm + 4 * msz;  c.X2a = sm + 5 * msz;
        c.X1b = sm + 6 * msz;  c.X2b = sm + 7 * msz; c.X3b = sm + 8 * msz;
        c.J0pc = J0p; c.J0pnew = J0pn;
        auto set_buf = [&](int b) {
            c.r = b ? lr1 : lr0; c.t = b ? lt1 : lt0; c.jp = jpb + b * n; c.jm = jmb + b * n;
        };

        // The surface as the "layer" below the last one, written in place into layer buffer b:
        // Lambertian (lambertian_surface.jl:41-95; only m = 0 has one) or a BRDF given by the Fourier
        // moments of R_surf (create_surface_layer!(::AbstractSurfaceType), rpv_surface.jl:51-97).
        T jps_sun = T(0);
        const bool has_surf = (m == 0) || (a.surf_rho != nullptr);
        auto fill_surface = [&](int b) {
            const T ts_end = a.p_tausum[(size_t)a.Nz * a.nS + p];
            const T att = Num<T>::exp_(-ts_end / a.mu0);
            const T* rho_m = a.surf_rho ? a.surf_rho + (size_t)m * n * n : nullptr;
            const T rho_l = T(2) * a.albedo;
            T* r = lbuf + (2 * b) * LB;
            T* sjp = jpb + b * n; T* sjm = jmb + b * n;
            for (int e = tid; e < n * n; e += blockDim.x) {
                const int i = e % n, j = e / n;
                const bool ii = (i % a.nStokes) == 0 && (j % a.nStokes) == 0;
                const T rho = rho_m ? rho_m[e] : (ii ? rho_l : T(0));
                r[i + j * ld] = rho * (a.mu[j] * a.wt[j]);
            }
            for (int i = tid; i < n; i += blockDim.x) {
                const T rho = rho_m ? rho_m[i + (size_t)a.istart * n] : (((i % a.nStokes) == 0) ? rho_l : T(0));
                sjp[i] = (i == a.istart) ? att : T(0);
                sjm[i] = a.mu0 * rho * att;
                if (a.sif0 && !rho_m && m == 0) sjm[i] += T(2) * a.sif0[(i % a.nStokes) + (size_t)a.nStokes * (a.F0_off + p)];
            }
            if (b) lr1 = r; else lr0 = r;
            jps_sun = att;
        };
        const int nsteps = a.Nz + (has_surf ? 1 : 0);     // step Nz = the surface
        const bool full = a.st_Rmp || a.st_Tpp;           // composite R-+ / T++ wanted (rt_run_streams)

        fetch_layer(0, p, mi, 0, c.RA, c.TC, c.J0pc, J0m);
        if (a.Nz > 1) fetch_layer(1, p, mi, 1, nullptr, nullptr, jpb + n, jmb + n);
        else if (has_surf) fill_surface(1);
        layer_ready(1);                       // layer 0 (element copies) and layer 1
        {
            const bool sc = (lflag[0] & 1) != 0;
            T* RA = c.RA; T* RB = c.RB; T* TC = c.TC; T* TD = c.TD;
            for (int e = tid; e < n * n; e += blockDim.x) {
                const int i = e % n, j = e / n;
                const uint32_t sg = dmask[i] ^ dmask[j];
                RB[i + j * ld] = sc ? Num<T>::flip(RA[i + j * ld], sg) : T(0);
                TD[i + j * ld] = Num<T>::flip(TC[i + j * ld], sg);
            }
            __syncthreads();
        }
        // one call site for the (large, inlined) interaction step: the composite's pointers stay in registers
        for (int iz = 1; iz < nsteps; ++iz) {
            const bool surf = (iz == a.Nz);
            const int fl = lflag[surf ? a.Nz - 1 : iz];
            const bool nosrc = surf || (fl & 1) == 0;     // the added layer has r+- = 0
            const int itf = fl >> 1;
            VSM_G0();
            // layer iz is complete in buffer iz&1 (waited + synchronised); start on iz+1
            const int nb = (iz + 1) & 1;
            if (iz + 1 < a.Nz) fetch_layer(iz + 1, p, mi, nb, nullptr, nullptr, jpb + nb * n, jmb + nb * n);
            else if (iz + 1 == a.Nz && has_surf) fill_surface(nb);
            set_buf(iz & 1);
            VSM_GTICK(5);
            if (itf == 3) add2_step11<T, NN, TM, GT>(c, nosrc, surf, full);
            else add2_step_other<T, NN, TM, GT>(c, itf, nosrc, surf);
            VSM_GTICK(6);
            layer_ready(nb);
            VSM_GTICK(7);
        }
        if (has_surf) {                       // hdr source: R_surf J0+ + j0-(surf)   (interaction_hdrf!)
            const T* J0pc = c.J0pc; const T* jm = c.jm;
            c.opC.template mv<false>(c.r, J0pc, [&](int i, T s) { hv[i] = s + jm[i]; });
        } else {
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

// ----------------------------------------------------------------------------
// k_interact_step -- L1 (vsm_rt_layer_*): ONE interaction!/copy_added_to_composite!
// (rt_kernel.jl:244-249) per spectral point with the CompositeLayer in global memory
// (the reference's (N,N,nSpec) arrays), staged through the arena.
// ----------------------------------------------------------------------------
template <typename T>
struct StepArgs {
    int N, nStokes, nS, ld, n_smem_mats;
    int iface, first, scatter;          // first: iz == 1 -> copy
    int neumann;                        // inverse mode of the specialised Ops (vsm_set_inverse_mode)
    const T* tau; const T* mu;          // non-scattering layer: t = diag(exp(-tau/mu))
    const T* add_r; const T* add_t; const T* add_jp; const T* add_jm;   // (N,N,nS)/(N,nS)
    T* Rmp; T* Rpm; T* Tpp; T* Tmm; T* J0p; T* J0m;                     // composite, in/out
    T* out_rpm; T* out_tmm;             // optional AddedLayer r+-, t-- outputs (may be null)
    T* out_rmp; T* out_tpp; T* out_jp; T* out_jm;   // non-scattering: AddedLayer fill (may be null)
    T* slab; size_t slab_stride;
    unsigned int* counter;
};

// NN > 0: compile-time N with the specialised (tensor-core) Ops of the column kernels, all nine matrices in shared memory
template <typename T, class Ops, int NN = 0>
__global__ void __launch_bounds__(256) k_interact_step(StepArgs<T> a, unsigned int total) {
    const int n = NN ? NN : a.N, ld = NN ? (NN | 1) : a.ld;
    const int msz = ld * n;
    T* sm = reinterpret_cast<T*>(vsm_smem);
    const int nsm = NN ? kAddMats : a.n_smem_mats;
    T* slab = a.slab + (size_t)blockIdx.x * a.slab_stride;
    auto mat = [&](int k) -> T* { return k < nsm ? sm + k * msz : slab + (size_t)(k - nsm) * msz; };
    T* vbase = sm + nsm * msz;
    T* J0p = vbase;         T* J0m = J0p + n;   T* jp = J0m + n;   T* jm = jp + n;
    T* v1 = jm + n;         T* hv = v1 + n;     T* vec = hv + n;
    T* part = vec + 5 * n;
    uint32_t* dmask = reinterpret_cast<uint32_t*>(part + 8 * n);
    int* piv = reinterpret_cast<int*>(dmask + n);
    __shared__ unsigned int sh_item;
    Ops op;
    if constexpr (NN == 0) { op.n = n; op.ld = ld; }
    op.vec = vec; op.part = part; op.piv = piv; op.dmask = dmask; op.tid = threadIdx.x; op.bar = 0;
    op.neumann = a.neumann;
    const size_t NN2 = (size_t)n * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dmask[i] = dmask_of(i, a.nStokes);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_item = atomicAdd(a.counter, 1u);
        __syncthreads();
        const unsigned int p = sh_item;
        if (p >= total) break;
        CompPtrs<T> c{mat(0), mat(1), mat(2), mat(3), mat(4), mat(5), mat(6), mat(7), mat(8)};
        const bool sc = a.scatter != 0;
        T* dr = a.first ? c.RA : c.r;
        T* dt = a.first ? c.TC : c.t;
        T* djp = a.first ? J0p : jp;
        T* djm = a.first ? J0m : jm;
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e % n, j = e / n;
            T rv, tv;
            if (sc) { rv = a.add_r[p * NN2 + e]; tv = a.add_t[p * NN2 + e]; }
            else { rv = T(0); tv = (i == j) ? Num<T>::exp_(-a.tau[p] / a.mu[i]) : T(0); }
            dr[i + j * ld] = rv;
            dt[i + j * ld] = tv;
            const uint32_t sg = dmask[i] ^ dmask[j];
            if (sc) {
                if (a.out_rpm) a.out_rpm[p * NN2 + e] = Num<T>::flip(rv, sg);
                if (a.out_tmm) a.out_tmm[p * NN2 + e] = Num<T>::flip(tv, sg);
            } else {          // zero_added_noscat! (rt_helpers.jl:174-180)
                if (a.out_rmp) a.out_rmp[p * NN2 + e] = T(0);
                if (a.out_rpm) a.out_rpm[p * NN2 + e] = T(0);
                if (a.out_tpp) a.out_tpp[p * NN2 + e] = tv;
                if (a.out_tmm) a.out_tmm[p * NN2 + e] = tv;
            }
            if (!a.first) {
                c.RA[i + j * ld] = a.Rmp[p * NN2 + e];
                c.RB[i + j * ld] = a.Rpm[p * NN2 + e];
                c.TC[i + j * ld] = a.Tpp[p * NN2 + e];
                c.TD[i + j * ld] = a.Tmm[p * NN2 + e];
            }
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            djp[i] = sc ? a.add_jp[(size_t)p * n + i] : T(0);
            djm[i] = sc ? a.add_jm[(size_t)p * n + i] : T(0);
            if (!sc && a.out_jm) a.out_jm[(size_t)p * n + i] = T(0);
            if (!a.first) { J0p[i] = a.J0p[(size_t)p * n + i]; J0m[i] = a.J0m[(size_t)p * n + i]; }
        }
        __syncthreads();
        if (a.first) {
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                const int i = e % n, j = e / n;
                const uint32_t sg = dmask[i] ^ dmask[j];
                c.RB[i + j * ld] = sc ? Num<T>::flip(c.RA[i + j * ld], sg) : T(0);
                c.TD[i + j * ld] = Num<T>::flip(c.TC[i + j * ld], sg);
            }
            __syncthreads();
        } else {
            interact<T, Ops>(op, a.iface, c, J0p, J0m, jp, jm, v1, !sc, false);
        }
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e % n, j = e / n;
            a.Rmp[p * NN2 + e] = c.RA[i + j * ld];
            a.Rpm[p * NN2 + e] = c.RB[i + j * ld];
            a.Tpp[p * NN2 + e] = c.TC[i + j * ld];
            a.Tmm[p * NN2 + e] = c.TD[i + j * ld];
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            a.J0p[(size_t)p * n + i] = J0p[i];
            a.J0m[(size_t)p * n + i] = J0m[i];
        }
    }
}

// out[o + nOut*p] = sum_m part[(m*nS + p)*nOut + o]   (m ascending, as the Fourier loop), then the Cox-Munk
// TMS correction R += mu0 exp(-tau_total/mu0) ss[o] (apply_ss_correction!, coxmunk_surface.jl:544-552) when given
template <typename T>
__global__ void k_reduce(int n_moments, size_t per_m, const T* __restrict__ pR,
                         const T* __restrict__ pT, const T* __restrict__ pH, T* __restrict__ R,
                         T* __restrict__ Tt, T* __restrict__ H, const T* __restrict__ ss,
                         const T* __restrict__ tau_total, T mu0, int nOut, int* __restrict__ nonfinite = nullptr) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_m) return;
    T r = T(0), t = T(0), h = T(0);
    for (int m = 0; m < n_moments; ++m) {
        r += pR[(size_t)m * per_m + e];
        t += pT[(size_t)m * per_m + e];
        if (H) h += pH[(size_t)m * per_m + e];
    }
    if (ss) r += (mu0 * Num<T>::exp_(-tau_total[e / nOut] / mu0)) * ss[e % nOut];
    // a singular (I - A) or an overflow anywhere upstream ends up here as Inf / NaN: counted per call
    if (nonfinite && !(Num<T>::abs_(r) <= T(1e300) && Num<T>::abs_(t) <= T(1e300))) atomicAdd(nonfinite, 1);
    R[e] = r;
    Tt[e] = t;
    if (H) H[e] = h;
}

}  // namespace vsm

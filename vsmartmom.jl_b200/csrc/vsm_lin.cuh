// vsm_lin.cuh -- LinMode (tangent-linear) column: forward R, T and the analytic Jacobians dR/dp, dT/dp for P
// parameters that act through the layer optical depths (Rayleigh / gas / aerosol) and the Lambertian albedo.
//
// Replaces, fused: rt_kernel!(::noRS, ..., lin) (CoreKernel/rt_kernel_lin.jl:50-180): the fused-chain-rule
// elemental! (elemental_lin.jl:77-206, kernels :456-712), doubling_allparams! (doubling_lin.jl:216-364), the lin
// interaction_helper!(::ScatteringInterface_11) (interaction_lin.jl:217-331), the Lambertian surface derivative
// (lambertian_surface_lin.jl) and the lin postprocessing_vza! (postprocessing_vza_lin.jl:18-49), driven like
// rt_run_lin.jl:102-326.
//
// Every quantity is a "dual": the forward value followed by its P derivatives, (1+P) consecutive matrices (or
// vectors) in the CTA's slab in global memory.  The forward algorithm is written once on duals:
//   C = A B            ->  C' = A' B + A B'                 (dual_mm: 1 + 2P products)
//   G = (I - A)^-1     ->  G' = G A' G                      (dual_inv: the forward inverse is reused for all P,
//                                                            doubling_lin.jl:43-53 / :268-273)
//   y = A x            ->  y' = A' x + A x'
// which reproduces the reference's hand-written rules term by term (12 products per parameter and doubling).
// Products run through the same Ops objects as the forward kernels: OpsStream (N = 60 / 75 / 100: DMMA, operands
// streamed from the L2-resident slab) or OpsGeneric (any N).
#pragma once
#include "vsm_column.cuh"

namespace vsm {

constexpr int kMaxParams = 16;

template <typename T>
struct LinArgs {
    int P;                        // number of parameters
    size_t msz;                   // elements per matrix in the slab / hand-off (ld * ncols)
    int vstride;                  // elements per vector slot (>= N)
    // k_props_lin outputs, [(p*Nz + iz)*nS + s] (tausum: (Nz+1) rows, row iz = above layer iz)
    const T* d_tau; const T* d_varpi; const T* d_tausum;
    const T* d_W;                 // [((p*nC + c)*Nz + iz)*nS + s]
    const T* dalbedo;             // [P] (device)
    // hand-off of the derivative layers: slot as the forward one, (1+P) matrices per slot
    T* add_r; T* add_t;           // [slot*(1+P) + q] * msz
    T* add_jp; T* add_jm;         // [slot*(1+P) + q] * N
    // per-moment partial Jacobians [((m*nS + s)*P + p)*nOut + o]
    T* part_dR; T* part_dT;
};

// ----------------------------------------------------------------------------
// k_props_lin: derivatives of the mixed layer properties (compEffectiveLayerProperties_lin.jl; with
// S = sum_c tau_c varpi_c: varpi = S/tau, W_c = tau_c varpi_c / S)
// ----------------------------------------------------------------------------
template <typename T>
struct PropsLinArgs {
    int nS, Nz, nAer, P;
    size_t ldspec;                // nSpec of the call (stride between layers, and between parameters x Nz)
    const T* tau_rayl; const T* tau_abs;          // offset to the chunk start
    const T* dtau_rayl; const T* dtau_abs;        // (nSpec, Nz, P) or null, offset to the chunk start
    T varpi_cab;
    const T* aer_tau_mod; const T* aer_varpi_mod; // as PropsArgs
    const T* daer_tau_mod;                        // [(p*Nz + iz)*nAer + c] or null
    T* d_tau; T* d_varpi; T* d_tausum; T* d_W;
};

template <typename T>
__global__ void k_props_lin(PropsLinArgs<T> a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.nS) return;
    const int nC = a.nAer + 1;
    for (int p = 0; p < a.P; ++p) {
        T dts = T(0);
        a.d_tausum[((size_t)p * (a.Nz + 1)) * a.nS + s] = dts;
        for (int iz = 0; iz < a.Nz; ++iz) {
            T tc[kMaxComp], vc[kMaxComp], dtc[kMaxComp];
            tc[0] = a.tau_rayl[s + (size_t)iz * a.ldspec];
            vc[0] = a.varpi_cab;
            dtc[0] = a.dtau_rayl ? a.dtau_rayl[s + ((size_t)p * a.Nz + iz) * a.ldspec] : T(0);
            for (int c = 0; c < a.nAer; ++c) {
                tc[c + 1] = a.aer_tau_mod[c + a.nAer * iz];
                vc[c + 1] = a.aer_varpi_mod[c];
                dtc[c + 1] = a.daer_tau_mod ? a.daer_tau_mod[((size_t)p * a.Nz + iz) * a.nAer + c] : T(0);
            }
            T tau = a.tau_abs[s + (size_t)iz * a.ldspec];
            T dtau = a.dtau_abs ? a.dtau_abs[s + ((size_t)p * a.Nz + iz) * a.ldspec] : T(0);
            T S = T(0), dS = T(0);
            for (int c = 0; c < nC; ++c) {
                tau += tc[c]; dtau += dtc[c];
                S = fma(tc[c], vc[c], S); dS = fma(dtc[c], vc[c], dS);
            }
            const T st = tau > T(0) ? tau : T(1), sS = S > T(0) ? S : T(1);
            const T varpi = S / st;
            const size_t o = ((size_t)p * a.Nz + iz) * a.nS + s;
            a.d_tau[o] = dtau;
            a.d_varpi[o] = (dS - varpi * dtau) / st;
            for (int c = 0; c < nC; ++c) {
                const T W = tc[c] * vc[c] / sS;
                a.d_W[(((size_t)p * nC + c) * a.Nz + iz) * a.nS + s] = (dtc[c] * vc[c] - dS * W) / sS;
            }
            dts += dtau;
            a.d_tausum[((size_t)p * (a.Nz + 1) + iz + 1) * a.nS + s] = dts;
        }
    }
}

// ----------------------------------------------------------------------------
// dual helpers
// ----------------------------------------------------------------------------
template <typename T>
struct Dual {
    T* f;                 // forward matrix; derivative q at f + (q+1)*msz
    size_t msz;
    __device__ __forceinline__ T* d(int q) const { return f + (size_t)(q + 1) * msz; }
};

// C = A B (+ Cin when ACC) on duals.  tmp: not needed (the second term of every derivative accumulates in place).
// KSIGN / csign: A D B with the result columns signed by D (the t-- = D t D products of interaction.jl:240,265).
template <typename T, class Ops, bool KSIGN, bool ACC>
__device__ __forceinline__ void dual_mm(const Ops& op, int P, const Dual<T>& A, const Dual<T>& B, const Dual<T>& C,
                                        bool a_const = false, bool b_const = false) {
    const int ld = op.LD();
    const uint32_t* dmask = op.dmask;
    auto put = [&](T* M, bool acc) {
        return [=](int i, int j, T v) {
            if (KSIGN) v = Num<T>::flip(v, dmask[j]);
            T* e = M + i + (size_t)j * ld;
            *e = acc ? *e + v : v;
        };
    };
    if constexpr (Ops::kIsStream) {
        // Streamed operands: the products of one dual product are known in advance and none of them writes an operand of
        // another (C is distinct from A and B), so every product ANNOUNCES the next one -- its first panels are in flight
        // while the current epilogue runs and it starts warm instead of refilling the cp.async ring from cold.
        const int per_q = (a_const ? 0 : 1) + (b_const ? 0 : 1);
        const int total = P * per_q + 1;
        auto operands = [&](int k, const T*& pa, const T*& pb) {
            if (k >= total - 1) { pa = A.f; pb = B.f; return; }
            const int q = k / per_q;
            const bool first_of_q = (k % per_q) == 0;
            if (first_of_q && !a_const) { pa = A.d(q); pb = B.f; }
            else { pa = A.f; pb = B.d(q); }
        };
        auto none = [](int, T) {};
        if (per_q == 0) {
            for (int q = 0; q < P; ++q)
                for (int e = op.tid; e < op.N() * op.N(); e += op.nthreads()) C.d(q)[(e % op.N()) + (size_t)(e / op.N()) * ld] = T(0);
            op.sync();
        }
        for (int k = 0; k < total; ++k) {
            const T *pa, *pb, *na = nullptr, *nb = nullptr;
            operands(k, pa, pb);
            if (k + 1 < total) operands(k + 1, na, nb);
            const bool last = (k == total - 1);
            T* dst = last ? C.f : C.d(k / per_q);
            const bool acc = last ? ACC : (ACC || (k % per_q) != 0);
            op.template mm_stream<KSIGN, 0>(pa, pb, put(dst, acc), pa, pa, none, none, na, nb);
            op.sync();
        }
    } else {
    for (int q = 0; q < P; ++q) {                         // derivatives first: they read the old forward C only if ACC
        bool first = !ACC;
        if (!a_const) { op.template mm<KSIGN>(A.d(q), B.f, put(C.d(q), !first)); first = false; op.sync(); }
        if (!b_const) { op.template mm<KSIGN>(A.f, B.d(q), put(C.d(q), !first)); first = false; op.sync(); }
        if (first) {                                      // both operands constant: C' = 0
            for (int e = op.tid; e < op.N() * op.N(); e += op.nthreads()) C.d(q)[(e % op.N()) + (size_t)(e / op.N()) * ld] = T(0);
            op.sync();
        }
    }
    op.template mm<KSIGN>(A.f, B.f, put(C.f, ACC));
    op.sync();
    }
}

// y = A x on duals (vectors: forward at v, derivative q at v + (q+1)*vs); epi(q, i, value) with q = -1 forward
template <typename T, class Ops, class Epi>
__device__ __forceinline__ void dual_mv(const Ops& op, int P, const Dual<T>& A, const T* x, int vs, T* tmp, Epi epi,
                                        bool a_const = false) {
    op.template mv<false>(A.f, x, [&](int i, T s) { epi(-1, i, s); });
    for (int q = 0; q < P; ++q) {
        if (a_const) {
            op.template mv<false>(A.f, x + (size_t)(q + 1) * vs, [&](int i, T s) { epi(q, i, s); });
        } else {
            op.template mv<false>(A.d(q), x, [&](int i, T s) { tmp[i] = s; });
            op.template mv<false>(A.f, x + (size_t)(q + 1) * vs, [&](int i, T s) { epi(q, i, tmp[i] + s); });
        }
    }
}

// G = (I - A B)^-1 on duals: forward by the Ops' adaptive inverse, G' = G (A' B + A B') G.  S1, S2: scratch matrices.
template <typename T, class Ops>
__device__ __forceinline__ void dual_inv_one_minus(const Ops& op, int P, const Dual<T>& A, const Dual<T>& B, const Dual<T>& G,
                                                   T* S1, T* S2) {
    const int ld = op.LD();
    op.inv_one_minus(A.f, B.f, G.f, S1);
    if constexpr (Ops::kIsStream) {
        // as in dual_mm: a product whose operands are complete is announced one product ahead (A'B -> AB', and the last
        // product of parameter q -> the first of q+1); G S1 and S2 G read what the previous product wrote and start cold
        auto none = [](int, T) {};
        for (int q = 0; q < P; ++q) {
            op.template mm_stream<false, 0>(A.d(q), B.f, [&](int i, int j, T v) { S1[i + (size_t)j * ld] = v; }, A.f, A.f, none, none,
                                            A.f, B.d(q));
            op.sync();
            op.template mm_stream<false, 0>(A.f, B.d(q), [&](int i, int j, T v) { S1[i + (size_t)j * ld] += v; }, A.f, A.f, none, none,
                                            nullptr, nullptr);
            op.sync();
            op.template mm_stream<false, 0>(G.f, S1, [&](int i, int j, T v) { S2[i + (size_t)j * ld] = v; }, A.f, A.f, none, none,
                                            nullptr, nullptr);
            op.sync();
            T* Gq = G.d(q);
            const bool more = q + 1 < P;
            op.template mm_stream<false, 0>(S2, G.f, [&](int i, int j, T v) { Gq[i + (size_t)j * ld] = v; }, A.f, A.f, none, none,
                                            more ? A.d(q + 1) : nullptr, more ? B.f : nullptr);
            op.sync();
        }
    } else {
    for (int q = 0; q < P; ++q) {
        op.template mm<false>(A.d(q), B.f, [&](int i, int j, T v) { S1[i + (size_t)j * ld] = v; });
        op.sync();
        op.template mm<false>(A.f, B.d(q), [&](int i, int j, T v) { S1[i + (size_t)j * ld] += v; });
        op.sync();
        op.template mm<false>(G.f, S1, [&](int i, int j, T v) { S2[i + (size_t)j * ld] = v; });
        op.sync();
        T* Gq = G.d(q);
        op.template mm<false>(S2, G.f, [&](int i, int j, T v) { Gq[i + (size_t)j * ld] = v; });
        op.sync();
    }
    }
}

// ----------------------------------------------------------------------------
// elemental layer and its derivatives (elemental_lin.jl:77-206 with lin_added_layer_all_params!): the oracle's
// formulation, one thread per element, P inner loop.  Writes r, t (duals) and j0+, j0- (dual vectors).
// ----------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void elemental_lin_fill(const ColArgs<T>& a, const LinArgs<T>& l, int n, int ld, int m, int nd, int iz,
                                                   int sp, T dtau, T varpi, T tausum, T scale, const T* Wc,
                                                   const T* Zpp, const T* Zmp, size_t zstride, const T* F0n,
                                                   const uint32_t* dmask, const Dual<T>& r, const Dual<T>& t,
                                                   T* jp, T* jm, int vs, T* cap /* 2*(1+P)*nStokes*n */) {
    const int nC = a.nC, P = l.P, nSt = a.nStokes;
    const T wsc = (m == 0) ? T(0.5) : T(0.25);
    const T eps = Num<T>::eps();
    T dd[kMaxParams], dw[kMaxParams];
    for (int q = 0; q < P; ++q) {
        const size_t o = ((size_t)q * a.Nz + iz) * a.nS + sp;
        dd[q] = l.d_tau[o] / scale;
        dw[q] = l.d_varpi[o];
    }
    // solar columns of Z and Z' are captured here for the sources: cap[((q*2 + pm)*nSt + s)*n + i], q = 0 forward
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e % n, j = e / n;
        T zcp[kMaxComp], zcm[kMaxComp];
        T Zp = T(0), Zm = T(0);
        for (int c = 0; c < nC; ++c) {
            zcp[c] = Zpp[(size_t)c * zstride + e];
            zcm[c] = Zmp[(size_t)c * zstride + e];
            const T w = (nC > 1) ? Wc[c] : T(1);
            Zp = fma(w, zcp[c], Zp);
            Zm = fma(w, zcm[c], Zm);
        }
        const T mui = a.mu[i], muj = a.mu[j];
        const T wj = a.wt[j] * wsc, wi = a.wt[i] * wsc;
        const bool active = wj > eps;
        const T cc = T(1) / mui + T(1) / muj;
        const T g1 = (muj / (mui + muj)) * wj;
        const T E1 = -Num<T>::expm1_(-dtau * cc), dE1 = cc * Num<T>::exp_(-dtau * cc);
        const T ei = Num<T>::exp_(-dtau / mui), ej = Num<T>::exp_(-dtau / muj);
        T rv = T(0), tv = T(0);
        T g2E2 = T(0), g2dE2 = T(0);
        const bool same = (mui == muj);
        if (active) {
            rv = varpi * Zm * g1 * E1;
            if (!same) {
                const T g2 = (muj / (mui - muj)) * wj;
                g2E2 = g2 * expdiff_neg<T>(dtau / mui, dtau / muj);
                g2dE2 = g2 * (-ei / mui + ej / muj);
                tv = varpi * Zp * g2E2;
            } else if (i == j) {
                tv = ei * (T(1) + varpi * Zp * (dtau / mui) * wi);
            } else {
                tv = ej * varpi * Zp * (dtau / mui) * wj;
            }
        } else {
            tv = (i == j) ? ei : T(0);
        }
        const uint32_t fl = (nd >= 1) ? dmask[i] : 0u;
        r.f[i + (size_t)j * ld] = Num<T>::flip(rv, fl);
        t.f[i + (size_t)j * ld] = tv;
        const bool sun = (j >= a.istart && j < a.istart + nSt);
        if (sun) {
            cap[((0 * 2 + 0) * nSt + (j - a.istart)) * n + i] = Zp;
            cap[((0 * 2 + 1) * nSt + (j - a.istart)) * n + i] = Zm;
        }
        for (int q = 0; q < P; ++q) {
            T dZp = T(0), dZm = T(0);
            if (nC > 1) {
                for (int c = 0; c < nC; ++c) {
                    const T dW = l.d_W[(((size_t)q * nC + c) * a.Nz + iz) * a.nS + sp];
                    dZp = fma(dW, zcp[c], dZp);
                    dZm = fma(dW, zcm[c], dZm);
                }
            }
            if (sun) {
                cap[(((q + 1) * 2 + 0) * nSt + (j - a.istart)) * n + i] = dZp;
                cap[(((q + 1) * 2 + 1) * nSt + (j - a.istart)) * n + i] = dZm;
            }
            T drv = T(0), dtv = T(0);
            if (active) {
                drv = (dw[q] * Zm + varpi * dZm) * (g1 * E1) + (varpi * Zm * g1 * dE1) * dd[q];
                const T wz = dw[q] * Zp + varpi * dZp;
                if (!same) {
                    dtv = wz * g2E2 + (varpi * Zp * g2dE2) * dd[q];
                } else if (i == j) {
                    dtv = (ei * (dtau / mui) * wi) * wz
                          + (-(T(1) / mui) * ei * (T(1) + varpi * Zp * (dtau / mui) * wi) + ei * varpi * Zp * wi / mui) * dd[q];
                } else {
                    dtv = (ej * (dtau / mui) * wj) * wz
                          + (-(T(1) / muj) * ej * varpi * Zp * (dtau / mui) * wj + ej * varpi * Zp * wj / mui) * dd[q];
                }
            } else if (i == j) {
                dtv = -(ei / mui) * dd[q];
            }
            r.d(q)[i + (size_t)j * ld] = Num<T>::flip(drv, fl);
            t.d(q)[i + (size_t)j * ld] = dtv;
        }
    }
    __syncthreads();
    // solar sources (elemental_lin.jl SFI kernel; oracle elemental_lin)
    const T mus = a.mu[a.istart];
    const T att = Num<T>::exp_(-tausum / mus);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const T mui = a.mu[i];
        const bool insun = (i >= a.istart && i < a.istart + nSt);
        const T ei = Num<T>::exp_(-dtau / mui);
        T fp, dfp;
        if (insun) {
            fp = (dtau / mui) * ei;
            dfp = ei / mui - (dtau / mui) * ei / mui;
        } else {
            const T gs2 = mus / (mui - mus);
            fp = gs2 * expdiff_neg<T>(dtau / mui, dtau / mus);
            dfp = gs2 * (-ei / mui + Num<T>::exp_(-dtau / mus) / mus);
        }
        const T cs = T(1) / mui + T(1) / mus;
        const T fm = (mus / (mui + mus)) * (-Num<T>::expm1_(-dtau * cs));
        const T dfm = (mus / (mui + mus)) * cs * Num<T>::exp_(-dtau * cs);
        T ZpI = T(0), ZmI = T(0);
        for (int s = 0; s < nSt; ++s) {
            ZpI = fma(cap[((0 * 2 + 0) * nSt + s) * n + i], F0n[s], ZpI);
            ZmI = fma(cap[((0 * 2 + 1) * nSt + s) * n + i], F0n[s], ZmI);
        }
        const T jp0 = wsc * varpi * ZpI * fp, jm0 = wsc * varpi * ZmI * fm;
        const uint32_t fl = (nd >= 1) ? dmask[i] : 0u;
        jp[i] = jp0 * att;
        jm[i] = Num<T>::flip(jm0 * att, fl);
        for (int q = 0; q < P; ++q) {
            T dZpI = T(0), dZmI = T(0);
            for (int s = 0; s < nSt; ++s) {
                dZpI = fma(cap[(((q + 1) * 2 + 0) * nSt + s) * n + i], F0n[s], dZpI);
                dZmI = fma(cap[(((q + 1) * 2 + 1) * nSt + s) * n + i], F0n[s], dZmI);
            }
            const T datt = -(att / mus) * l.d_tausum[((size_t)q * (a.Nz + 1) + iz) * a.nS + sp];
            const T djp0 = wsc * ((dw[q] * ZpI + varpi * dZpI) * fp + varpi * ZpI * dfp * dd[q]);
            const T djm0 = wsc * ((dw[q] * ZmI + varpi * dZmI) * fm + varpi * ZmI * dfm * dd[q]);
            jp[(size_t)(q + 1) * vs + i] = djp0 * att + jp0 * datt;
            jm[(size_t)(q + 1) * vs + i] = Num<T>::flip(djm0 * att + jm0 * datt, fl);
        }
    }
    __syncthreads();
}

// slab layout (in matrices of msz): k_layer_lin uses 7 duals + 2 singles, k_adding_lin 13 duals + 2 singles
constexpr int kLinLayerDuals = 7, kLinAddDuals = 13, kLinSingles = 2;
// dual vectors per CTA (each (1+P)*vstride): layer: jp jm j1p j1m am ap + tmp ; adding: J0p J0m J0pn jp jm v1 v2 + tmp
constexpr int kLinLayerVecs = 7, kLinAddVecs = 8;

template <typename T, class Ops, int NN, int NTHR, int MINB = 1>
__global__ void __launch_bounds__(NTHR, MINB) k_layer_lin(ColArgs<T> a, LinArgs<T> l, unsigned int total) {
    const int n = NN ? NN : a.N;
    T* smv = reinterpret_cast<T*>(vsm_smem);
    Ops op;
    int ld;
    T* part; T* vec; uint32_t* dmask; int* piv; T* Wc;
    // shared memory: [Ops ring (stream only)] part(8) vec(5) Wc dmask piv  -- sized by the host as for the forward kernels
    if constexpr (Ops::kIsStream) {
        using CF = typename Ops::Cfg;
        ld = CF::LD;
        T* vb = smv + CF::NST * CF::STAGE;
        part = vb; vec = part + 8 * CF::NP; Wc = vec + 5 * CF::NP;
        dmask = reinterpret_cast<uint32_t*>(Wc + kMaxComp); piv = reinterpret_cast<int*>(dmask + CF::NP);
        op.stages = smv;
        for (int i = threadIdx.x; i < CF::NP; i += blockDim.x) dmask[i] = (i < n) ? dmask_of(i, a.nStokes) : 0u;
    } else {
        ld = a.ld;
        part = smv; vec = part + 8 * n; Wc = vec + 5 * n;
        dmask = reinterpret_cast<uint32_t*>(Wc + kMaxComp); piv = reinterpret_cast<int*>(dmask + n);
        op.n = n; op.ld = ld;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dmask[i] = dmask_of(i, a.nStokes);
    }
    op.vec = vec; op.part = part; op.piv = piv; op.dmask = dmask; op.tid = threadIdx.x; op.bar = 0; op.neumann = a.neumann;
    const int P = l.P, vs = l.vstride;
    const size_t msz = l.msz, dsz = (size_t)(1 + P) * msz;
    T* slab = a.slab + (size_t)blockIdx.x * a.slab_stride;
    Dual<T> r{slab, msz}, t{slab + dsz, msz}, G{slab + 2 * dsz, msz}, X2{slab + 3 * dsz, msz}, W{slab + 4 * dsz, msz},
        TN{slab + 5 * dsz, msz}, A1{slab + 6 * dsz, msz};
    T* S1 = slab + 7 * dsz; T* S2 = S1 + msz;
    T* vbase = S2 + msz;                                        // dual vectors live in the slab too
    const size_t dv = (size_t)(1 + P) * vs;
    T* jp = vbase; T* jm = jp + dv; T* j1p = jm + dv; T* j1m = j1p + dv; T* am = j1m + dv; T* ap = am + dv; T* tmpv = ap + dv;
    T* cap = tmpv + dv;                                         // 2*(1+P)*nStokes*n
    __shared__ unsigned int sh_item;
    unsigned int nextv = 0;
    if (threadIdx.x == 0) nextv = atomicAdd(a.counter, 1u);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_item = nextv;
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= total) break;
        if (threadIdx.x == 0) nextv = atomicAdd(a.counter, 1u);
        const int sp = item % a.nS;
        const unsigned int rest = item / a.nS;
        const int mi = rest % a.Mb;
        const int si = rest / a.Mb;
        const int iz = a.layer_order[si];
        const int m = a.m0 + mi;
        const int nd = a.ndoubl[iz];
        const size_t pi = (size_t)iz * a.nS + sp;
        const T tau = a.p_tau[pi], varpi = a.p_varpi[pi], tausum = a.p_tausum[pi];
        const T scale = (T)exp2((double)nd);
        const T dtau = tau / scale;
        T expk = Num<T>::exp_(-dtau / a.mu0);
        if (threadIdx.x < a.nC) Wc[threadIdx.x] = a.p_W[((size_t)threadIdx.x * a.Nz + iz) * a.nS + sp];
        __syncthreads();
        const size_t zstride = (size_t)n * n;
        const T* Zpp = a.Zpp + (size_t)(m - a.mz0) * a.nC * zstride;
        const T* Zmp = a.Zmp + (size_t)(m - a.mz0) * a.nC * zstride;
        elemental_lin_fill<T>(a, l, n, ld, m, nd, iz, sp, dtau, varpi, tausum, scale, Wc, Zpp, Zmp, zstride,
                              a.F0 + (a.F0_off + sp) * a.nStokes, dmask, r, t, jp, jm, vs, cap);
        // d(expk)/dp = -expk/mu0 * d(dtau)/dp   (doubling_lin.jl:257-259)
        T dexpk[kMaxParams];
        for (int q = 0; q < P; ++q) dexpk[q] = -(expk / a.mu0) * (l.d_tau[((size_t)q * a.Nz + iz) * a.nS + sp] / scale);
        Dual<T> rr = r, tt = t, TNN = TN;
        for (int it = 0; it < nd; ++it) {
            // G = (I - r r)^-1 and G'   (:268-273)
            dual_inv_one_minus<T, Ops>(op, P, rr, rr, G, S1, S2);
            // X2 = t G   (tt of the reference, :274)
            dual_mm<T, Ops, false, false>(op, P, tt, G, X2);
            // sources (:283-306)
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                j1p[i] = jp[i] * expk; j1m[i] = jm[i] * expk;
                for (int q = 0; q < P; ++q) {
                    j1p[(q + 1) * vs + i] = jp[(q + 1) * vs + i] * expk + jp[i] * dexpk[q];
                    j1m[(q + 1) * vs + i] = jm[(q + 1) * vs + i] * expk + jm[i] * dexpk[q];
                }
            }
            __syncthreads();
            // am = j1m + r jp ; ap = jp + r j1m  (and derivatives)
            dual_mv<T, Ops>(op, P, rr, jp, vs, tmpv, [&](int q, int i, T s) { am[(q + 1) * vs + i] = j1m[(q + 1) * vs + i] + s; });
            dual_mv<T, Ops>(op, P, rr, j1m, vs, tmpv, [&](int q, int i, T s) { ap[(q + 1) * vs + i] = jp[(q + 1) * vs + i] + s; });
            // jm += X2 am ; jp = j1p + X2 ap
            dual_mv<T, Ops>(op, P, X2, am, vs, tmpv, [&](int q, int i, T s) { jm[(q + 1) * vs + i] += s; });
            dual_mv<T, Ops>(op, P, X2, ap, vs, tmpv, [&](int q, int i, T s) { jp[(q + 1) * vs + i] = j1p[(q + 1) * vs + i] + s; });
            for (int q = 0; q < P; ++q) dexpk[q] = T(2) * expk * dexpk[q];      // (:294)
            expk = expk * expk;
            // W = r t ; r += X2 W ; t = X2 t   (:310-324)
            dual_mm<T, Ops, false, false>(op, P, rr, tt, W);
            dual_mm<T, Ops, false, true>(op, P, X2, W, rr);
            dual_mm<T, Ops, false, false>(op, P, X2, tt, TNN);
            Dual<T> sw = tt; tt = TNN; TNN = sw;
        }
        // apply_D_matrix! / apply_D_matrix_SFI! (doubling_lin.jl:330-334) on the way out to the hand-off
        const size_t slot = ((size_t)mi * a.n_sched_layers + si) * a.nS + sp;
        for (int q = 0; q <= P; ++q) {
            T* gr = l.add_r + (slot * (1 + P) + q) * msz;
            T* gt = l.add_t + (slot * (1 + P) + q) * msz;
            const T* sr = rr.f + (size_t)q * msz; const T* st = tt.f + (size_t)q * msz;
            for (size_t e = threadIdx.x; e < msz; e += blockDim.x) {
                const int i = (int)(e % ld), j = (int)(e / ld);
                const bool in = (i < n && j < n);
                T rv = in ? sr[e] : T(0);
                if (nd >= 1 && in) rv = Num<T>::flip(rv, dmask[i]);
                gr[e] = rv;
                gt[e] = in ? st[e] : T(0);
            }
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                T vm = jm[(size_t)q * vs + i];
                if (nd >= 1) vm = Num<T>::flip(vm, dmask[i]);
                l.add_jp[(slot * (1 + P) + q) * n + i] = jp[(size_t)q * vs + i];
                l.add_jm[(slot * (1 + P) + q) * n + i] = vm;
            }
        }
    }
}

// ----------------------------------------------------------------------------
// k_adding_lin: the interaction chain on duals for one (point, moment): copy_added_to_composite!, Nz-1 x lin
// interaction_helper!(::ScatteringInterface_11), the surface (Lambertian with its albedo derivative, or BRDF moments),
// and the lin postprocessing_vza!.  Only all-scattering columns (every interface _11): the host rejects others.
// ----------------------------------------------------------------------------
template <typename T, class Ops, int NN, int NTHR, int MINB = 1>
__global__ void __launch_bounds__(NTHR, MINB) k_adding_lin(ColArgs<T> a, LinArgs<T> l, unsigned int total) {
    const int n = NN ? NN : a.N;
    T* smv = reinterpret_cast<T*>(vsm_smem);
    Ops op;
    int ld;
    T* part; T* vec; uint32_t* dmask; int* piv;
    if constexpr (Ops::kIsStream) {
        using CF = typename Ops::Cfg;
        ld = CF::LD;
        T* vb = smv + CF::NST * CF::STAGE;
        part = vb; vec = part + 8 * CF::NP;
        dmask = reinterpret_cast<uint32_t*>(vec + 5 * CF::NP + kMaxComp); piv = reinterpret_cast<int*>(dmask + CF::NP);
        op.stages = smv;
        for (int i = threadIdx.x; i < CF::NP; i += blockDim.x) dmask[i] = (i < n) ? dmask_of(i, a.nStokes) : 0u;
    } else {
        ld = a.ld;
        part = smv; vec = part + 8 * n;
        dmask = reinterpret_cast<uint32_t*>(vec + 5 * n + kMaxComp); piv = reinterpret_cast<int*>(dmask + n);
        op.n = n; op.ld = ld;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dmask[i] = dmask_of(i, a.nStokes);
    }
    op.vec = vec; op.part = part; op.piv = piv; op.dmask = dmask; op.tid = threadIdx.x; op.bar = 0; op.neumann = a.neumann;
    const int P = l.P, vs = l.vstride, tid = threadIdx.x;
    const size_t msz = l.msz, dsz = (size_t)(1 + P) * msz;
    T* slab = a.slab + (size_t)blockIdx.x * a.slab_stride;
    // (slots 0, 2, 6 and 10 of the slab held the composite R-+ / T++ and their scratch; they are not advanced, see below)
    Dual<T> RB{slab + dsz, msz}, TD{slab + 3 * dsz, msz}, G{slab + 4 * dsz, msz},
        T01{slab + 5 * dsz, msz}, T21{slab + 7 * dsz, msz}, Y{slab + 8 * dsz, msz},
        N1{slab + 9 * dsz, msz}, N3{slab + 11 * dsz, msz}, SR{slab + 12 * dsz, msz};
    T* S1 = slab + 13 * dsz; T* S2 = S1 + msz;
    const size_t dv = (size_t)(1 + P) * vs;
    T* J0p = S2 + msz; T* J0m = J0p + dv; T* J0pn = J0m + dv; T* jp = J0pn + dv; T* jm = jp + dv;
    T* v1 = jm + dv; T* v2 = v1 + dv; T* tmpv = v2 + dv;
    const int nOut = a.nVZA * a.nStokes;
    auto at = [ld](T* M, int i, int j) -> T& { return M[i + (size_t)j * ld]; };
    __shared__ unsigned int sh_item;
    unsigned int nextv = 0;
    if (tid == 0) nextv = atomicAdd(a.counter, 1u);
    for (;;) {
        __syncthreads();
        if (tid == 0) sh_item = nextv;
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= total) break;
        if (tid == 0) nextv = atomicAdd(a.counter, 1u);
        const int sp = item % a.nS;
        const int mi = item / a.nS;
        const int m = a.m0 + mi;
        Dual<T> cRB = RB, cTD = TD, nTD = N1, nRB = N3;
        T* cJ0p = J0p; T* cJ0pn = J0pn;
        const bool has_surf = (m == 0) || (a.surf_rho != nullptr);
        const int nsteps = a.Nz + (has_surf ? 1 : 0);
        // ---- layer 0: copy_added_to_composite! on duals
        {
            const size_t slot = ((size_t)mi * a.n_sched_layers + a.sched_idx[0]) * a.nS + sp;
            for (int q = 0; q <= P; ++q) {
                const T* gr = l.add_r + (slot * (1 + P) + q) * msz;
                const T* gt = l.add_t + (slot * (1 + P) + q) * msz;
                T* rb = cRB.f + (size_t)q * msz; T* td = cTD.f + (size_t)q * msz;
                for (int e = tid; e < n * n; e += blockDim.x) {
                    const int i = e % n, j = e / n;
                    const size_t o = i + (size_t)j * ld;
                    const uint32_t sg = dmask[i] ^ dmask[j];
                    const T rv = gr[o], tv = gt[o];
                    rb[o] = Num<T>::flip(rv, sg); td[o] = Num<T>::flip(tv, sg);
                }
                for (int i = tid; i < n; i += blockDim.x) {
                    cJ0p[(size_t)q * vs + i] = l.add_jp[(slot * (1 + P) + q) * n + i];
                    J0m[(size_t)q * vs + i] = l.add_jm[(slot * (1 + P) + q) * n + i];
                }
            }
            __syncthreads();
        }
        for (int iz = 1; iz < nsteps; ++iz) {
            const bool surf = (iz == a.Nz);
            Dual<T> r, t;
            if (!surf) {
                const size_t slot = ((size_t)mi * a.n_sched_layers + a.sched_idx[iz]) * a.nS + sp;
                r = Dual<T>{l.add_r + slot * (1 + P) * msz, msz};
                t = Dual<T>{l.add_t + slot * (1 + P) * msz, msz};
                for (int q = 0; q <= P; ++q)
                    for (int i = tid; i < n; i += blockDim.x) {
                        jp[(size_t)q * vs + i] = l.add_jp[(slot * (1 + P) + q) * n + i];
                        jm[(size_t)q * vs + i] = l.add_jm[(slot * (1 + P) + q) * n + i];
                    }
            } else {
                // surface dual (lambertian_surface_lin.jl; BRDF moments have no parameter here): r = rho mu w,
                // j0+ = att e_sun, j0- = mu0 rho att ; att' = -att/mu0 tausum', rho' = dalbedo (Lambertian only)
                const T ts_end = a.p_tausum[(size_t)a.Nz * a.nS + sp];
                const T att = Num<T>::exp_(-ts_end / a.mu0);
                const T* rho_m = a.surf_rho ? a.surf_rho + (size_t)m * n * n : nullptr;
                const T rho_l = T(2) * a.albedo;
                r = SR; t = Dual<T>{nullptr, msz};
                for (int q = 0; q <= P; ++q) {
                    const T drho = (q == 0) ? rho_l : T(2) * l.dalbedo[q - 1];
                    T* rq = SR.f + (size_t)q * msz;
                    for (int e = tid; e < n * n; e += blockDim.x) {
                        const int i = e % n, j = e / n;
                        const bool ii = (i % a.nStokes) == 0 && (j % a.nStokes) == 0;
                        T rho;
                        if (rho_m) rho = (q == 0) ? rho_m[e] : T(0);
                        else rho = ii ? drho : T(0);
                        rq[i + (size_t)j * ld] = rho * (a.mu[j] * a.wt[j]);
                    }
                    const T datt = (q == 0) ? att : -(att / a.mu0) * l.d_tausum[((size_t)(q - 1) * (a.Nz + 1) + a.Nz) * a.nS + sp];
                    for (int i = tid; i < n; i += blockDim.x) {
                        const T rho0 = rho_m ? rho_m[i + (size_t)a.istart * n] : (((i % a.nStokes) == 0) ? rho_l : T(0));
                        const T drho_i = rho_m ? T(0) : (((i % a.nStokes) == 0) ? drho : T(0));
                        jp[(size_t)q * vs + i] = (i == a.istart) ? datt : T(0);
                        jm[(size_t)q * vs + i] = (q == 0) ? a.mu0 * rho0 * att : a.mu0 * (drho_i * att + rho0 * datt);
                    }
                }
            }
            __syncthreads();
            // ---- minus half (interaction_lin.jl:238-267)
            dual_inv_one_minus<T, Ops>(op, P, r, cRB, G, S1, S2);                    // G1 = (I - r R+-)^-1
            dual_mm<T, Ops, false, false>(op, P, cTD, G, T01);                       // T01 = T-- G1
            dual_mv<T, Ops>(op, P, r, cJ0p, vs, tmpv, [&](int q, int i, T s) { v1[(q + 1) * vs + i] = s + jm[(q + 1) * vs + i]; });
            dual_mv<T, Ops>(op, P, T01, v1, vs, tmpv, [&](int q, int i, T s) { J0m[(q + 1) * vs + i] += s; });
            // (after the surface step only J0+- are read: its matrix updates are skipped.  The composite R-+ and
            //  T++ (:252-255, :296) enter no J0+- / T-- / R+- of any later step, i.e. no LinMode output: not advanced.)
            if (!surf) dual_mm<T, Ops, true, false>(op, P, T01, t, nTD);             // T--' = T01 t--
            // ---- plus half (:280-306)
            dual_inv_one_minus<T, Ops>(op, P, cRB, r, G, S1, S2);                    // G2 = (I - R+- r)^-1
            Dual<T> T21u = G;                                                        // (surface: t = I)
            if (!surf) { dual_mm<T, Ops, false, false>(op, P, t, G, T21); T21u = T21; }   // T21 = t G2
            dual_mv<T, Ops>(op, P, cRB, jm, vs, tmpv, [&](int q, int i, T s) { v2[(q + 1) * vs + i] = cJ0p[(q + 1) * vs + i] + s; });
            dual_mv<T, Ops>(op, P, T21u, v2, vs, tmpv, [&](int q, int i, T s) { cJ0pn[(q + 1) * vs + i] = jp[(q + 1) * vs + i] + s; });
            if (!surf) {
                dual_mm<T, Ops, false, false>(op, P, T21u, cRB, Y);                  // Y = T21 R+-
                dual_mm<T, Ops, true, false>(op, P, Y, t, nRB);                      // R+-' = r+- + Y t--
                for (int q = 0; q <= P; ++q) {
                    T* dst = nRB.f + (size_t)q * msz; const T* src = r.f + (size_t)q * msz;
                    for (int e = tid; e < n * n; e += blockDim.x) {
                        const int i = e % n, j = e / n;
                        dst[i + (size_t)j * ld] += Num<T>::flip(src[i + (size_t)j * ld], dmask[i] ^ dmask[j]);
                    }
                }
                __syncthreads();
                Dual<T> sw;
                sw = cTD; cTD = nTD; nTD = sw;
                sw = cRB; cRB = nRB; nRB = sw;
            }
            { T* sw = cJ0p; cJ0p = cJ0pn; cJ0pn = sw; }
        }
        // ---- lin postprocessing_vza! into per-moment partials
        const size_t pbase = ((size_t)m * a.nS + sp) * nOut;
        for (int o = tid; o < nOut; o += blockDim.x) {
            const int iv = o % a.nVZA, s = o / a.nVZA;
            const T w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
            const int row = a.vza_row[iv] + s;
            a.part_R[pbase + o] = w * J0m[row];
            a.part_T[pbase + o] = w * cJ0p[row];
            for (int q = 0; q < P; ++q) {
                const size_t pb = (((size_t)m * a.nS + sp) * P + q) * nOut;
                l.part_dR[pb + o] = w * J0m[(size_t)(q + 1) * vs + row];
                l.part_dT[pb + o] = w * cJ0p[(size_t)(q + 1) * vs + row];
            }
        }
    }
}

// dR[o + nOut*(s + nSpec*q)] = sum_m part_dR[((m*nS + s)*P + q)*nOut + o]   (m ascending)
template <typename T>
__global__ void k_reduce_lin(int n_moments, int nS, int P, int nOut, size_t nspec_total, size_t s0,
                             const T* __restrict__ pR, const T* __restrict__ pT, T* __restrict__ dR, T* __restrict__ dT) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_m = (size_t)nS * P * nOut;
    if (e >= per_m) return;
    const int o = (int)(e % nOut);
    const int q = (int)((e / nOut) % P);
    const size_t s = e / ((size_t)nOut * P);
    T r = T(0), t = T(0);
    for (int m = 0; m < n_moments; ++m) {
        r += pR[(size_t)m * per_m + e];
        t += pT[(size_t)m * per_m + e];
    }
    const size_t dst = o + (size_t)nOut * (s0 + s + nspec_total * q);
    dR[dst] = r;
    dT[dst] = t;
}

}  // namespace vsm

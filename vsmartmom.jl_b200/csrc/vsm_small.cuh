// vsm_small.cuh -- the noRS column for small matrices (N <= 16: quickstart N = 4, the file default Stokes_I / 8 streams
// N = 13, the Raman test scene N = 15), fp64.  One WARP owns one (spectral point, Fourier moment): the whole chain
// elemental! -> doubling! -> interaction! over every layer, the surface step, interaction_hdrf! and postprocessing_vza!
// (rt_run.jl:383-517, rt_kernel.jl:175-250) runs on nine zero-padded 16 x 16 matrices in the warp's private shared-memory
// arena, every product on the fp64 tensor cores (DMMA, fragments in the k order of vsm_raman.cuh), no CTA barrier and no
// hand-off through HBM.  The CTA-per-item kernels (k_layer / k_adding) leave most of a CTA idle at these sizes.
#pragma once
#include "vsm_column.cuh"
#include "vsm_raman.cuh"

namespace vsm {
namespace small {
using namespace vsm::rrs;

constexpr int kMats = 9, kVecs = 13;
template <int NP, int WPC>
constexpr size_t smem_bytes() { return (size_t)WPC * (kMats * Geo<NP>::MS + kVecs * NP) * sizeof(double) + 64; }

template <int NP, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_small(ColArgs<double> a, unsigned int total) {
    using G = Geo<NP>;
    constexpr int LDS = G::LDS, MS = G::MS;
    extern __shared__ __align__(16) unsigned char small_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * WPC + warp;
    if (item >= total) return;
    const int p = item % a.nS;
    const int m = a.m0 + item / a.nS;
    double* base = reinterpret_cast<double*>(small_smem) + (size_t)warp * (kMats * MS + kVecs * NP);
    double* r = base; double* t = r + MS; double* Gm = t + MS; double* TG = Gm + MS; double* S = TG + MS;
    double* Rmp = S + MS; double* Rpm = Rmp + MS; double* Tpp = Rpm + MS; double* Tmm = Tpp + MS;
    double* jp = Tmm + MS; double* jm = jp + NP; double* j1p = jm + NP; double* j1m = j1p + NP;
    double* a1 = j1m + NP; double* a2 = a1 + NP; double* colk = a2 + NP; double* sg = colk + NP;
    double* J0p = sg + NP; double* J0m = J0p + NP; double* hv = J0m + NP; double* mu = hv + NP;
    int* piv = reinterpret_cast<int*>(mu + NP);
    const int n = a.N;
    for (int i = lane; i < NP; i += 32) {
        sg[i] = (i < n && (i % a.nStokes) >= 2) ? -1.0 : 1.0;
        mu[i] = (i < n) ? a.mu[i] : 1.0;
    }
    __syncwarp();
    double c[NP / 8][NP / 8][2];
    double y1[NP / 16], y2[NP / 16];
    const SM<NP> sr{r}, st{t}, sG{Gm}, sTG{TG}, sS{S}, sRmp{Rmp}, sRpm{Rpm}, sTpp{Tpp}, sTmm{Tmm};
    const Conj<SM<NP>> stmm{st, sg};
    const size_t zs = (size_t)n * n;
    const double* Zpp = a.Zpp + (size_t)(m - a.mz0) * a.nC * zs;
    const double* Zmp = a.Zmp + (size_t)(m - a.mz0) * a.nC * zs;
    const double* F0 = a.F0 + (a.F0_off + p) * a.nStokes;
    const double wsc = (m == 0) ? 0.5 : 0.25;

    // one ScatteringInterface step of the composite (in shared memory) with the added layer (r, t, jp, jm); rpm0: r+- = 0
    auto interact = [&](int itf, bool rpm0) {
        if (itf == 3) {                                              // interaction.jl:207-266
            czero<NP>(c);
            wmm<NP>(c, sr, sRpm, lane);
            winv_one_minus<NP>(c, Gm, S, TG, n, a.neumann, colk, piv, lane);
            czero<NP>(c);                                            // T01 = T-- inv1
            wmm<NP>(c, sTmm, sG, lane);
            wepi<NP>(c, [&](int i, int j, double v) { TG[i + j * LDS] = v; }, lane);
            wmv<NP>(y1, sr, SV{J0p}, lane);
            if (lane < 16) for (int b = 0; b < NP / 16; ++b) a1[16 * b + lane] = jm[16 * b + lane] + y1[b];
            __syncwarp();
            wmv<NP>(y2, sTG, SV{a1}, lane);
            if (lane < 16) for (int b = 0; b < NP / 16; ++b) J0m[16 * b + lane] += y2[b];
            // (R-+ += T01 r T++ (:237) and T++ = T21 T++ (:262) are not advanced: the composite R-+ / T++ enter no
            //  J0+- / T-- / R+- of any later step, and this kernel never exports them)
            czero<NP>(c);                                            // T-- = T01 t--
            wmm<NP>(c, sTG, stmm, lane);
            wepi<NP>(c, [&](int i, int j, double v) { Tmm[i + j * LDS] = v; }, lane);
            czero<NP>(c);                                            // inv2 = (I - R+- r)^-1
            wmm<NP>(c, sRpm, sr, lane);
            __syncwarp();
            winv_one_minus<NP>(c, Gm, S, TG, n, a.neumann, colk, piv, lane);
            czero<NP>(c);                                            // T21 = t inv2
            wmm<NP>(c, st, sG, lane);
            wepi<NP>(c, [&](int i, int j, double v) { TG[i + j * LDS] = v; }, lane);
            wmv<NP>(y1, sRpm, SV{jm}, lane);
            if (lane < 16) for (int b = 0; b < NP / 16; ++b) a2[16 * b + lane] = J0p[16 * b + lane] + y1[b];
            __syncwarp();
            wmv<NP>(y2, sTG, SV{a2}, lane);
            if (lane < 16) for (int b = 0; b < NP / 16; ++b) J0p[16 * b + lane] = jp[16 * b + lane] + y2[b];
            czero<NP>(c);                                            // R+- = r+- + (T21 R+-) t--
            wmm<NP>(c, sTG, sRpm, lane);
            wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);
            __syncwarp();
            czero<NP>(c);
            wmm<NP>(c, sS, stmm, lane);
            __syncwarp();
            wepi<NP>(c, [&](int i, int j, double v) { Rpm[i + j * LDS] = (rpm0 ? 0.0 : r[i + j * LDS] * sg[i] * sg[j]) + v; }, lane);
            __syncwarp();
            return;
        }
        if (itf == 1) {                                              // interaction.jl:96-118
            wmv<NP>(y1, sr, SV{J0p}, lane);
            if (lane < 16) for (int b = 0; b < NP / 16; ++b) a1[16 * b + lane] = jm[16 * b + lane] + y1[b];
            __syncwarp();
            wmv<NP>(y1, sTmm, SV{a1}, lane);
            wmv<NP>(y2, st, SV{J0p}, lane);
            __syncwarp();
            if (lane < 16)
                for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; J0m[i] += y1[b]; J0p[i] = jp[i] + y2[b]; }
            czero<NP>(c);                                            // R-+ = (T-- r) T++
            wmm<NP>(c, sTmm, sr, lane);
            wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);
            __syncwarp();
            czero<NP>(c);
            wmm<NP>(c, sS, sTpp, lane);
            wepi<NP>(c, [&](int i, int j, double v) { Rmp[i + j * LDS] = v; Rpm[i + j * LDS] = rpm0 ? 0.0 : r[i + j * LDS] * sg[i] * sg[j]; }, lane);
            czero<NP>(c);                                            // T++ = t T++
            wmm<NP>(c, st, sTpp, lane);
            __syncwarp();
            wepi<NP>(c, [&](int i, int j, double v) { Tpp[i + j * LDS] = v; }, lane);
            czero<NP>(c);                                            // T-- = T-- t--
            wmm<NP>(c, sTmm, stmm, lane);
            __syncwarp();
            wepi<NP>(c, [&](int i, int j, double v) { Tmm[i + j * LDS] = v; }, lane);
            __syncwarp();
            return;
        }
        if (itf == 2) {                                              // interaction.jl:141-162
            wmv<NP>(y1, sRpm, SV{jm}, lane);
            if (lane < 16) for (int b = 0; b < NP / 16; ++b) a1[16 * b + lane] = J0p[16 * b + lane] + y1[b];
            __syncwarp();
            wmv<NP>(y1, st, SV{a1}, lane);
            wmv<NP>(y2, sTmm, SV{jm}, lane);
            __syncwarp();
            if (lane < 16)
                for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; J0p[i] = jp[i] + y1[b]; J0m[i] += y2[b]; }
            czero<NP>(c);                                            // T++ = t T++
            wmm<NP>(c, st, sTpp, lane);
            __syncwarp();
            wepi<NP>(c, [&](int i, int j, double v) { Tpp[i + j * LDS] = v; }, lane);
            czero<NP>(c);                                            // T-- = T-- t--
            wmm<NP>(c, sTmm, stmm, lane);
            __syncwarp();
            wepi<NP>(c, [&](int i, int j, double v) { Tmm[i + j * LDS] = v; }, lane);
            czero<NP>(c);                                            // R+- = (t R+-) t--
            wmm<NP>(c, st, sRpm, lane);
            wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);
            __syncwarp();
            czero<NP>(c);
            wmm<NP>(c, sS, stmm, lane);
            __syncwarp();
            wepi<NP>(c, [&](int i, int j, double v) { Rpm[i + j * LDS] = v; }, lane);
            __syncwarp();
            return;
        }
        // ScatteringInterface_00 (interaction.jl:52-72)
        wmv<NP>(y1, st, SV{J0p}, lane);
        wmv<NP>(y2, sTmm, SV{jm}, lane);
        __syncwarp();
        if (lane < 16)
            for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; J0p[i] = jp[i] + y1[b]; J0m[i] += y2[b]; }
        czero<NP>(c);                                                // T-- = t-- T--
        wmm<NP>(c, stmm, sTmm, lane);
        __syncwarp();
        wepi<NP>(c, [&](int i, int j, double v) { Tmm[i + j * LDS] = v; }, lane);
        czero<NP>(c);                                                // T++ = t T++
        wmm<NP>(c, st, sTpp, lane);
        __syncwarp();
        wepi<NP>(c, [&](int i, int j, double v) { Tpp[i + j * LDS] = v; }, lane);
        __syncwarp();
    };

    for (int iz = 0; iz < a.Nz; ++iz) {
        const bool sc = a.scatter[iz] != 0;
        const size_t pi = (size_t)iz * a.nS + p;
        const double tau = a.p_tau[pi];
        if (sc) {
            const int nd = a.ndoubl[iz];
            const double varpi = a.p_varpi[pi], tausum = a.p_tausum[pi];
            const double dtau = tau / exp2((double)nd);
            double expk = exp(-dtau / a.mu0);
            // ---- elemental! (elemental.jl:289-392, 403-422); per-row transcendentals as in vsm_raman.cuh:
            //      u_i = expm1(-dτ/μ_i);  -expm1(-dτ(1/μ_i+1/μ_j)) = -(u_i + u_j + u_i u_j),  exp(-dτ/μ_i) = 1 + u_i
            for (int i = lane; i < NP; i += 32) { a1[i] = expm1(-dtau / mu[i]); jp[i] = 0.0; jm[i] = 0.0; }
            __syncwarp();
            for (int e = lane; e < NP * NP; e += 32) {
                const int i = e % NP, j = e / NP;
                double rv = 0.0, tv = 0.0;
                if (i < n && j < n) {
                    double zpp = 0.0, zmp = 0.0;
                    for (int cc = 0; cc < a.nC; ++cc) {
                        const double w = a.p_W[((size_t)cc * a.Nz + iz) * a.nS + p];
                        zpp = fma(w, Zpp[cc * zs + i + (size_t)j * n], zpp);
                        zmp = fma(w, Zmp[cc * zs + i + (size_t)j * n], zmp);
                    }
                    const double mui = mu[i], muj = mu[j], wct = a.wt[j] * wsc;
                    const double ui = a1[i], uj = a1[j];
                    if (wct > 2.220446049250313e-16) {
                        rv = varpi * zmp * (muj / (mui + muj)) * wct * (-(ui + uj + ui * uj));
                        if (mui == muj) {
                            if (i == j) tv = (1.0 + ui) * (1.0 + varpi * zpp * (dtau / mui) * wct);
                            else tv = (1.0 + uj) * (varpi * zpp * (dtau / mui) * wct);
                        } else {
                            const double x = dtau / mui, y = dtau / muj;
                            const double ed = (fabs(x - y) < 1e-3 * fmax(x, y)) ? expdiff_neg<double>(x, y) : (ui - uj);
                            tv = varpi * zpp * (muj / (mui - muj)) * wct * ed;
                        }
                    } else {
                        tv = (i == j) ? (1.0 + ui) : 0.0;
                    }
                    if (nd >= 1) rv *= sg[i];
                }
                r[i + j * LDS] = rv;
                t[i + j * LDS] = tv;
            }
            {   // get_elem_rt_SFI! (elemental.jl:348-392)
                const double mus = mu[a.istart], us = a1[a.istart];
                const double att = exp(-tausum / mus);
                for (int i = lane; i < n; i += 32) {
                    double zpI = 0.0, zmI = 0.0;
                    for (int s = 0; s < a.nStokes; ++s) {
                        double zpp = 0.0, zmp = 0.0;
                        const int j = a.istart + s;
                        for (int cc = 0; cc < a.nC; ++cc) {
                            const double w = a.p_W[((size_t)cc * a.Nz + iz) * a.nS + p];
                            zpp = fma(w, Zpp[cc * zs + i + (size_t)j * n], zpp);
                            zmp = fma(w, Zmp[cc * zs + i + (size_t)j * n], zmp);
                        }
                        zpI = fma(zpp, F0[s], zpI);
                        zmI = fma(zmp, F0[s], zmI);
                    }
                    const double mui = mu[i], ui = a1[i];
                    double vp;
                    if (i >= a.istart && i < a.istart + a.nStokes) vp = wsc * varpi * zpI * (dtau / mui) * (1.0 + ui);
                    else {
                        const double x = dtau / mui, y = dtau / mus;
                        const double ed = (fabs(x - y) < 1e-3 * fmax(x, y)) ? expdiff_neg<double>(x, y) : (ui - us);
                        vp = wsc * varpi * zpI * (mus / (mui - mus)) * ed;
                    }
                    double vm = wsc * varpi * zmI * (mus / (mui + mus)) * (-(ui + us + ui * us));
                    vp *= att;
                    vm *= att;
                    if (nd >= 1) vm *= sg[i];
                    jp[i] = vp;
                    jm[i] = vm;
                }
            }
            __syncwarp();
            // ---- doubling! (doubling.jl:38-99, rt_helpers.jl:102-166)
            for (int l = 0; l < nd; ++l) {
                czero<NP>(c);
                wmm<NP>(c, sr, sr, lane);
                winv_one_minus<NP>(c, Gm, S, TG, n, a.neumann, colk, piv, lane);
                czero<NP>(c);                                        // tt = t gp
                wmm<NP>(c, st, sG, lane);
                wepi<NP>(c, [&](int i, int j, double v) { TG[i + j * LDS] = v; }, lane);
                for (int i = lane; i < NP; i += 32) { j1p[i] = jp[i] * expk; j1m[i] = jm[i] * expk; }
                __syncwarp();
                wmv<NP>(y1, sr, SV{jp}, lane);
                wmv<NP>(y2, sr, SV{j1m}, lane);
                if (lane < 16)
                    for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; a1[i] = j1m[i] + y1[b]; a2[i] = jp[i] + y2[b]; }
                __syncwarp();
                wmv<NP>(y1, sTG, SV{a1}, lane);
                wmv<NP>(y2, sTG, SV{a2}, lane);
                if (lane < 16)
                    for (int b = 0; b < NP / 16; ++b) { const int i = 16 * b + lane; jm[i] += y1[b]; jp[i] = j1p[i] + y2[b]; }
                czero<NP>(c);                                        // r += (tt r) t
                wmm<NP>(c, sTG, sr, lane);
                wepi<NP>(c, [&](int i, int j, double v) { S[i + j * LDS] = v; }, lane);
                __syncwarp();
                czero<NP>(c);
                wmm<NP>(c, sS, st, lane);
                wepi<NP>(c, [&](int i, int j, double v) { r[i + j * LDS] += v; }, lane);
                czero<NP>(c);                                        // t = tt t
                wmm<NP>(c, sTG, st, lane);
                __syncwarp();
                wepi<NP>(c, [&](int i, int j, double v) { t[i + j * LDS] = v; }, lane);
                __syncwarp();
                expk = expk * expk;
            }
            if (nd >= 1) {                                           // apply_D_matrix! / apply_D_matrix_SFI!
                for (int e = lane; e < NP * NP; e += 32) { const int i = e % NP, j = e / NP; r[i + j * LDS] *= sg[i]; }
                for (int i = lane; i < NP; i += 32) jm[i] *= sg[i];
            }
        } else {
            // zero_added_noscat! (rt_helpers.jl:174-180): r = 0, t = diag(exp(-τ/μ)), j = 0
            for (int e = lane; e < NP * NP; e += 32) {
                const int i = e % NP, j = e / NP;
                r[i + j * LDS] = 0.0;
                t[i + j * LDS] = (i == j && i < n) ? exp(-tau / mu[i]) : 0.0;
            }
            for (int i = lane; i < NP; i += 32) { jp[i] = 0.0; jm[i] = 0.0; }
        }
        __syncwarp();
        if (iz == 0) {                                               // copy_added_to_composite!
            for (int e = lane; e < NP * NP; e += 32) {
                const int i = e % NP, j = e / NP;
                const double rv = r[i + j * LDS], tv = t[i + j * LDS], s2 = sg[i] * sg[j];
                Rmp[i + j * LDS] = rv; Rpm[i + j * LDS] = sc ? rv * s2 : 0.0; Tpp[i + j * LDS] = tv; Tmm[i + j * LDS] = tv * s2;
            }
            for (int i = lane; i < NP; i += 32) { J0p[i] = jp[i]; J0m[i] = jm[i]; }
            __syncwarp();
        } else {
            interact(a.iface[iz], !sc);
        }
    }
    // ---- surface (lambertian_surface.jl:41-95 / rpv_surface.jl:51-97) + last interaction (rt_run.jl:459-465)
    const double ts_end = a.p_tausum[(size_t)a.Nz * a.nS + p];
    double jps_sun = 0.0;
    if (m == 0 || a.surf_rho) {
        const double att = exp(-ts_end / a.mu0);
        const double* rho_m = a.surf_rho ? a.surf_rho + (size_t)m * n * n : nullptr;
        const double rho_l = 2.0 * a.albedo;
        for (int e = lane; e < NP * NP; e += 32) {
            const int i = e % NP, j = e / NP;
            double rv = 0.0;
            if (i < n && j < n) {
                const bool ii = (i % a.nStokes) == 0 && (j % a.nStokes) == 0;
                const double rho = rho_m ? rho_m[i + (size_t)j * n] : (ii ? rho_l : 0.0);
                rv = rho * (mu[j] * a.wt[j]);
            }
            r[i + j * LDS] = rv;
            t[i + j * LDS] = (i == j && i < n) ? 1.0 : 0.0;
        }
        for (int i = lane; i < NP; i += 32) {
            double vp = 0.0, vm = 0.0;
            if (i < n) {
                const double rho = rho_m ? rho_m[i + (size_t)a.istart * n] : (((i % a.nStokes) == 0) ? rho_l : 0.0);
                vp = (i == a.istart) ? att : 0.0;
                vm = a.mu0 * rho * att;
                if (a.sif0 && !rho_m && m == 0) vm += 2.0 * a.sif0[(i % a.nStokes) + (size_t)a.nStokes * (a.F0_off + p)];
            }
            jp[i] = vp;
            jm[i] = vm;
        }
        jps_sun = att;
        __syncwarp();
        interact(a.iface[a.Nz - 1], true);
        wmv<NP>(y1, sr, SV{J0p}, lane);                              // interaction_hdrf!: hdr_J0- = r-+ J0+ + j0-
        if (lane < 16) for (int b = 0; b < NP / 16; ++b) hv[16 * b + lane] = y1[b] + jm[16 * b + lane];
    } else {
        for (int i = lane; i < NP; i += 32) hv[i] = 0.0;
    }
    __syncwarp();
    // ---- postprocessing_vza! into the per-moment partials (summed in Fourier-loop order by k_reduce)
    const int nOut = a.nVZA * a.nStokes;
    const size_t pbase = ((size_t)m * a.nS + p) * nOut;
    for (int o = lane; o < nOut; o += 32) {
        const int iv = o % a.nVZA, s = o / a.nVZA;
        const double w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
        const int row = a.vza_row[iv] + s;
        a.part_R[pbase + o] = w * J0m[row];
        a.part_T[pbase + o] = w * J0p[row];
        a.part_H[pbase + o] = w * hv[row];
    }
    if (m == 0 && a.bhr_uw && lane < a.nStokes) {
        const int s = lane;
        double up = 0.0, dw = 0.0;
        for (int j = s; j < n; j += a.nStokes) {
            up += hv[j] * a.wt[j] * mu[j];
            dw += J0p[j] * a.wt[j] * mu[j];
        }
        a.bhr_uw[s + (size_t)a.nStokes * p] = up;
        a.bhr_dw[s + (size_t)a.nStokes * p] = dw + jps_sun * mu[a.istart];
    }
}

}  // namespace small
}  // namespace vsm

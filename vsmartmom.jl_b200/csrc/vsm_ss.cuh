// vsm_ss.cuh -- single-scattering driver: rt_run_ss(::noRS, model, iBand) (rt_run.jl:593-755) with
// rt_kernel_ss!(::noRS{FT}) (CoreKernel/rt_kernel_ss.jl:156-236: elemental! of the whole layer with ndoubl = 0, no
// doubling) and interaction_ss! (CoreKernel/interaction_ss.jl:14-42).  Only the solar source vectors are carried:
// O(N nStokes) work per (point, layer, moment) -- an HBM-bound kernel (inputs tau_rayl / tau_abs, outputs R/T).
// One warp per (spectral point, Fourier moment), lanes over the stream rows.
#pragma once
#include "vsm_column.cuh"

namespace vsm {

template <typename T>
__global__ void __launch_bounds__(128) k_ss(ColArgs<T> a, unsigned int total, T* __restrict__ hem_R, T* __restrict__ hem_T) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * 4 + warp;
    if (item >= total) return;
    const int p = item % a.nS;
    const int m = item / a.nS;
    const int n = a.N;
    const size_t zs = (size_t)n * n;
    const T* Zpp = a.Zpp + (size_t)m * a.nC * zs;
    const T* Zmp = a.Zmp + (size_t)m * a.nC * zs;
    const T* F0 = a.F0 + (a.F0_off + p) * a.nStokes;
    const T wsc = (m == 0) ? T(0.5) : T(0.25);
    const T mus = a.mu[a.istart];
    const int nOut = a.nVZA * a.nStokes;
    const size_t pbase = ((size_t)m * a.nS + p) * nOut;
    T hr = T(0), ht = T(0);
    for (int i = lane; i < n; i += 32) {
        const T mui = a.mu[i];
        T J0p = T(0), J0m = T(0);
        for (int iz = 0; iz < a.Nz; ++iz) {
            const size_t pi = (size_t)iz * a.nS + p;
            const T tau = a.p_tau[pi], varpi = a.p_varpi[pi], tausum = a.p_tausum[pi];
            // get_elem_rt_SFI! (elemental.jl:348-392) with dτ = τ
            T zpI = T(0), zmI = T(0);
            for (int s = 0; s < a.nStokes; ++s) {
                T zpp = T(0), zmp = T(0);
                const int j = a.istart + s;
                for (int cc = 0; cc < a.nC; ++cc) {
                    const T w = a.p_W[((size_t)cc * a.Nz + iz) * a.nS + p];
                    zpp = fma(w, Zpp[cc * zs + i + (size_t)j * n], zpp);
                    zmp = fma(w, Zmp[cc * zs + i + (size_t)j * n], zmp);
                }
                zpI = fma(zpp, F0[s], zpI);
                zmI = fma(zmp, F0[s], zmI);
            }
            T jp;
            if (i >= a.istart && i < a.istart + a.nStokes) jp = wsc * varpi * zpI * (tau / mui) * Num<T>::exp_(-tau / mui);
            else jp = wsc * varpi * zpI * (mus / (mui - mus)) * expdiff_neg<T>(tau / mui, tau / mus);
            T jm = wsc * varpi * zmI * (mus / (mui + mus)) * (-Num<T>::expm1_(-tau * ((T(1) / mui) + (T(1) / mus))));
            const T att = Num<T>::exp_(-tausum / mus);
            jp *= att;
            jm *= att;
            if (iz == 0) { J0p = jp; J0m = jm; }
            else {                                        // get_interaction_ss! (interaction_ss.jl:14-19)
                J0p = J0p * Num<T>::exp_(-tau / mui) + jp;
                J0m = J0m + jm * Num<T>::exp_(-tausum / mui);
            }
        }
        // create_surface_layer! sources + interaction_ss! with τ_surf = 0 (rt_run.jl:699-721)
        const T ts_end = a.p_tausum[(size_t)a.Nz * a.nS + p];
        if (m == 0 || a.surf_rho) {
            const T att = Num<T>::exp_(-ts_end / a.mu0);
            const T rho = a.surf_rho ? a.surf_rho[(size_t)m * zs + i + (size_t)a.istart * n]
                                     : (((i % a.nStokes) == 0) ? T(2) * a.albedo : T(0));
            J0p += (i == a.istart) ? att : T(0);
            J0m += a.mu0 * rho * att * Num<T>::exp_(-ts_end / mui);
        }
        // postprocessing_vza!: this lane owns row i
        for (int iv = 0; iv < a.nVZA; ++iv) {
            const int s = i - a.vza_row[iv];
            if (s >= 0 && s < a.nStokes) {
                const T w = a.vza_w[s + a.nStokes * (iv + (size_t)a.nVZA * m)];
                a.part_R[pbase + iv + a.nVZA * s] = w * J0m;
                a.part_T[pbase + iv + a.nVZA * s] = w * J0p;
            }
        }
        if (m == 0 && (i % a.nStokes) == 0) { hr += J0m * mui * a.wt[i]; ht += J0p * mui * a.wt[i]; }
    }
    if (m == 0 && hem_R) {                                 // rt_run.jl:737-748
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { hr += __shfl_xor_sync(0xffffffffu, hr, o); ht += __shfl_xor_sync(0xffffffffu, ht, o); }
        if (lane == 0) { hem_R[p] = hr; hem_T[p] = ht; }
    }
}

}  // namespace vsm

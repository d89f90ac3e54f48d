// vsm_raman_args.h -- host-visible part of the rotational-Raman path (vsm_raman.cuh): record geometry, the kernel
// argument block and the per-size launch descriptor (kernels are instantiated in vsm_inst_raman.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace vsm {
namespace rrs {

template <int NP>
struct Geo {
    static constexpr int NT = NP / 8;          // 8x8 accumulator tiles per side
    static constexpr int LDS = NP + 1;         // shared leading dimension (== 1 mod 16 for NP = 16, 48)
    static constexpr int MS = LDS * NP;        // doubles per shared matrix
    static constexpr int MG = NP * NP;         // doubles per global (record) matrix, ld = NP
    static constexpr int VB = NP / 16;         // 16-row blocks of a warp mat-vec
    // record sizes (doubles)
    static constexpr int REC0 = 5 * MG + 5 * NP;   // level, donor side:  r t GT GR GRT | tmp1 tmp2 J0p J1m expk
    static constexpr int REC1 = 2 * MG;            // level, own side:    r' TG'  (stored transposed: A-side operands)
    static constexpr int IREC0 = 8 * MG + 4 * NP;  // interaction, donor: Rpm Tpp r tmm E1 E2 E3 E4 | v1 v2 J0p j0m
    static constexpr int IREC1 = 4 * MG;           // interaction, own:   T01' T21' r' Rpm'  (transposed)
    static constexpr int CE = 4 * MG + 2 * NP;     // elastic composite:  Rmp Rpm Tpp Tmm | J0p J0m
    static constexpr int CI = 4 * MG + 2 * NP;     // unit composite:     ieRmp ieRpm ieTpp ieTmm | ieJ0p ieJ0m
};

struct Args {
    int N, nStokes, Nz, nC, nVZA, nM;
    int nBand;               // spectral points of the band (n0 must lie in [0, nBand))
    int e0, nE;              // elastic points computed here: band indices [e0, e0+nE)
    int n1lo, n1cnt;         // inelastic owner range (band indices)
    int nRaman;
    int istart;              // 0-based first row of the solar stream
    int neumann;             // 1: adaptive Neumann / Gauss-Jordan inverse in the elastic column, 0: elimination only
    int order;               // unit enumeration: 0 = dn-major (n1 fastest), 1 = donor-major (dn fastest)
    int u_lo, u_n0cnt;       // donor-major enumeration: n0 in [u_lo, u_lo + u_n0cnt)
    const int2* units;       // order 1: the valid (n1, dn) units of the owner range sorted by donor wavelength (host-built), or null
    double mu0;
    const double* mu; const double* wt; const double* sgn;     // [NP] (sgn: +1 I,Q rows / -1 U,V rows; pad +1)
    const int* shift; const double* wR;                        // [nRaman]
    const int* ndoubl;                                         // [Nz]
    // k_props outputs over the elastic range, layout [iz*nE + p]; W [(c*Nz+iz)*nE + p]; fsc [iz*nE + p]
    const double* p_tau; const double* p_varpi; const double* p_tausum; const double* p_W; const double* fsc;
    const double* Zpp; const double* Zmp;        // (N,N,nC,nM)
    const double* ZppR; const double* ZmpR;      // (N,N,nM)   Raman Z moments (computeRamanZλ!)
    const double* F0;                            // (nStokes, nBand)
    const double* surf_rho;                      // (N,N,nM) or null (no reflecting surface at this moment)
    int lambertian;                              // 1: surface exists only at m == 0 (lambertian_surface.jl:41-95)
    double* cE_in; double* cE_out;               // [p][CE]
    double* rec0; double* rec1;                  // [l][p][REC0|REC1]
    double* irec0; double* irec1;                // [p][IREC0|IREC1]
    double* cI_in; double* cI_out;               // [dn*n1cnt + (n1-n1lo)][CI]
    double* uout;                                // [dn*n1cnt + (n1-n1lo)][2*nOut]: ieJ0- | ieJ0+ at the VZA rows, weighted
    const int* vza_row; const double* vza_w;     // [nVZA] ; (nStokes,nVZA,nM)
    double* R_SFI; double* T_SFI;                // (nVZA,nStokes,n1cnt), accumulated over m
};


struct Desc {
    int NP, WPC;
    size_t smem_elastic, smem_inel;
    size_t rec0, rec1, irec0, irec1, ce, ci;      // record sizes in doubles
    const void* k_elastic; const void* k_inel;
    cudaError_t (*launch_elastic)(cudaStream_t, const Args&, int m, int iz, int mode);
    cudaError_t (*launch_inel)(cudaStream_t, const Args&, int m, int iz, int mode, unsigned int total);
};
const Desc* desc(int N);          // nullptr when N is not covered (N > 48)
cudaError_t launch_fscatt(cudaStream_t, int nE, int Nz, int nAer, size_t ldspec, const double* tau_rayl,
                          const double* aer_tau_mod, double* fsc);
// argument block of the Raman single-scattering kernel (k_ss_rrs)
struct SsArgs {
    int N, nStokes, Nz, nVZA, nBand, nRaman, istart;
    const double* mu;                            // [N]
    const int* shift; const double* wR;          // [nRaman]
    const double* p_tau; const double* p_tausum; // [iz*nBand + n]
    const double* fsc;                           // [iz*nBand + n]
    const double* ZppR; const double* ZmpR;      // (N,N,nM)
    const double* F0;                            // (nStokes, nBand)
    const int* vza_row; const double* vza_w;     // [nVZA] ; (nStokes,nVZA,nM)
    double* ieR; double* ieT;                    // (nVZA,nStokes,nBand), accumulated over m
};
cudaError_t launch_ss_rrs(cudaStream_t, const SsArgs&, int m, unsigned int nBand);
cudaError_t launch_reduce(cudaStream_t, const Args&, double* ieR, double* ieT, int* nonfinite);

}  // namespace rrs
}  // namespace vsm

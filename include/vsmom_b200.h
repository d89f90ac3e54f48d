/*
 * vsmom_b200.h -- C-ABI of the B200-native CoreRT matrix-operator core.
 *
 * Drop-in boundary for the reference's CUDA package extension
 * (vSmartMOM.jl: ext/vSmartMOMCUDAExt.jl, ext/gpu_batched_cuda.jl) and for the
 * per-layer / per-column hot loop of rt_run (src/CoreRT/rt_run.jl:383-517).
 * A Julia `ccall` stub for every entry point is shown in INTEGRATION.md.
 *
 * Conventions
 *  - All arrays are column-major exactly as Julia lays them out; the spectral
 *    index n is the slowest dimension of every (N,N,nSpec) / (N,1,nSpec) array.
 *  - Pointers documented "device" are CUDA device pointers owned by the caller
 *    (CuArray memory); the library only borrows them for the duration of a call.
 *    Pointers documented "host" are ordinary host memory.
 *  - Every function returns 0 on success, non-zero on failure; the message is
 *    available from vsm_last_error().  Nothing throws across the ABI.
 *  - Work is enqueued on the cudaStream_t given to vsm_create(); entry points
 *    return after enqueueing unless stated otherwise (the reference calls
 *    CUDA.synchronize() after each op; the Julia glue does the same).
 *  - There is NO CPU fallback: without a CUDA device vsm_create() fails.
 */
#ifndef VSMOM_B200_H
#define VSMOM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSM_OK                 0
#define VSM_ERR_INVALID_ARG    1
#define VSM_ERR_CUDA           2
#define VSM_ERR_UNSUPPORTED    3
#define VSM_ERR_NOMEM          4

/* ScatteringInterface tags (src/CoreRT/types.jl AbstractScatteringInterface;
 * chosen by get_scattering_interface, src/CoreRT/tools/rt_helper_functions.jl:15-33) */
#define VSM_IFACE_00 0
#define VSM_IFACE_01 1
#define VSM_IFACE_10 2
#define VSM_IFACE_11 3

/* surface kinds handled natively by vsm_rt_column_* */
#define VSM_SURFACE_LAMBERTIAN_SCALAR 0   /* src/CoreRT/Surfaces/lambertian_surface.jl:41-95 */
#define VSM_SURFACE_BRDF_FOURIER      1   /* create_surface_layer!(::AbstractSurfaceType), rpv_surface.jl:51-97:
                                            the caller supplies the Fourier moments of R_surf */

typedef struct vsm_ctx vsm_ctx;   /* opaque; one per (device, host task) */

/* -- context -- replaces vSmartMOMCUDAExt.__init__ (ext/vSmartMOMCUDAExt.jl:36-66) */
int  vsm_version(void);
int  vsm_create(vsm_ctx** out, int device, void* cuda_stream /* cudaStream_t or NULL */);
int  vsm_destroy(vsm_ctx* ctx);
const char* vsm_last_error(vsm_ctx* ctx);       /* ctx may be NULL: last create error */
int  vsm_synchronize(vsm_ctx* ctx);             /* Architectures._sync_gpu[] (ext/vSmartMOMCUDAExt.jl:55) */
/* cap (bytes) on the library-owned scratch used by vsm_rt_column_*; 0 = default */
int  vsm_set_workspace_limit(vsm_ctx* ctx, uint64_t bytes);
/* how (I - A)^-1 of doubling (rt_helpers.jl:103-104) and interaction (interaction.jl:220-222,
 * :248-249) is evaluated by the specialised fp64 kernels.  1 (default): the Frobenius norm nu of
 * A is measured; if nu^K/(1-nu) < 1e-20 for K in {2,3,4,8} the Neumann series / product form
 * (I+A)(I+A^2)(I+A^4) truncated at order K is summed on the tensor cores (bit-level equivalent to
 * the elimination at fp64 precision), otherwise blocked Gauss-Jordan.  0: always Gauss-Jordan
 * (the reference's getrf/getri route, cpu_batched.jl:32-37). */
int  vsm_set_inverse_mode(vsm_ctx* ctx, int mode);
/* number of kernels this context has launched since creation (bench bookkeeping) */
uint64_t vsm_kernel_launches(vsm_ctx* ctx);
/* roofline denominator: measured register-resident FMA throughput (TFLOP/s) of this
 * device; kind 0 = fp64 FMA, 1 = fp32 FMA.  Blocks until done (~50 ms).            */
int  vsm_measure_fma_peak(vsm_ctx* ctx, int kind, double* tflops);

/* -- L0: batched dense algebra -- same semantics as ext/gpu_batched_cuda.jl ----
 * batched_mul  (gpu_batched_cuda.jl:208-218):  C[:,:,k] = A[:,:,k] * B[:,:,k]
 *   A: (M,K,batch)  B: (K,Ncols,batch)  C: (M,Ncols,batch), all device, dense.
 * batch_inv!   (gpu_batched_cuda.jl:97-182):   X[:,:,k] = inv(A[:,:,k])
 *   A is overwritten (as getrf does); info[k] = 0 or the 1-based index of a
 *   zero pivot (device int array, may be NULL -- the reference never reads it).
 * batch_solve! (gpu_batched_cuda.jl:72-94):    X[:,:,k] = A[:,:,k] \ B[:,:,k]   */
int vsm_batched_mul_f64(vsm_ctx*, double* C, const double* A, const double* B,
                        int M, int K, int Ncols, int batch);
int vsm_batched_mul_f32(vsm_ctx*, float* C, const float* A, const float* B,
                        int M, int K, int Ncols, int batch);
int vsm_batch_inv_f64(vsm_ctx*, double* X, double* A, int N, int batch, int* info);
int vsm_batch_inv_f32(vsm_ctx*, float* X, float* A, int N, int batch, int* info);
int vsm_batch_solve_f64(vsm_ctx*, double* X, double* A, const double* B, int N, int Ncols, int batch);
int vsm_batch_solve_f32(vsm_ctx*, float* X, float* A, const float* B, int N, int Ncols, int batch);

/* -- L1: one layer, one Fourier moment -- drop-in for
 * rt_kernel!(::noRS, ...) src/CoreRT/CoreKernel/rt_kernel.jl:175-250
 * (elemental! + doubling! + interaction!/copy_added_to_composite!).
 * All array pointers are device pointers.                                      */
typedef struct vsm_layer_desc {
    int32_t struct_size;     /* sizeof(vsm_layer_desc) */
    int32_t N;               /* NquadN = Nquad*nStokes */
    int32_t nStokes;         /* pol_type.n */
    int32_t nSpec;
    int32_t m;               /* Fourier moment */
    int32_t imu0;            /* 1-based iμ₀ (stream index) */
    int32_t ndoubl;          /* from get_dtau_ndoubl (rt_kernel.jl:266-287) on the FULL band */
    int32_t scatter;         /* maximum(τ.*ϖ) > 2eps(FT) on the FULL band (rt_kernel.jl:197) */
    int32_t iface;           /* VSM_IFACE_* for this layer */
    int32_t iz;              /* 1 => copy_added_to_composite! instead of interaction! */
    int32_t z_is_spectral;   /* size(Z,3) > 1 */
    int32_t reserved;
    double  mu0;             /* quad_points.μ₀ */
    const void* tau;         /* [nSpec]  total layer optical depth τ */
    const void* varpi;       /* [nSpec]  ϖ */
    const void* Zpp;         /* (N,N,nSpec) or (N,N,1) */
    const void* Zmp;
    const void* tau_sum;     /* [nSpec]  τ above the layer */
    const void* F0;          /* (nStokes,nSpec) */
    const void* mu;          /* [N] qp_μN */
    const void* wt;          /* [N] wt_μN */
    /* CompositeLayer (in/out) src/CoreRT/types.jl:155-174 */
    void *Rmp, *Rpm, *Tpp, *Tmm;   /* (N,N,nSpec) */
    void *J0p, *J0m;               /* (N,1,nSpec) */
    /* AddedLayer (out; any may be NULL => kept on-chip only) types.jl:196-230 */
    void *rmp, *rpm, *tpp, *tmm;
    void *j0p, *j0m;
} vsm_layer_desc;
int vsm_rt_layer_f64(vsm_ctx*, const vsm_layer_desc*);
int vsm_rt_layer_f32(vsm_ctx*, const vsm_layer_desc*);

/* -- L2: whole column, all Fourier moments, device epilogue -- drop-in for the
 * body of rt_run(::noRS, model, iBand) src/CoreRT/rt_run.jl:383-517 including
 * constructCoreOpticalProperties/expandOpticalProperties
 * (LayerOpticalProperties/compEffectiveLayerProperties.jl:11-117), the
 * Lambertian create_surface_layer!, interaction_hdrf! and postprocessing_vza!.
 * "host" arrays are small and always double; "device" arrays are f32/f64
 * according to the entry point.                                               */
typedef struct vsm_column_desc {
    int32_t struct_size;     /* sizeof(vsm_column_desc) */
    int32_t N;               /* NquadN */
    int32_t nStokes;
    int32_t nSpec;           /* spectral points in THIS call (a shard of the band) */
    int32_t Nz;
    int32_t nAer;
    int32_t nVZA;
    int32_t n_moments;       /* m_max + 1 */
    int32_t imu0;            /* 1-based iμ₀ */
    int32_t surface_kind;    /* VSM_SURFACE_* */
    double  mu0;
    double  albedo;          /* LambertianSurfaceScalar.albedo */
    double  varpi_cabannes;  /* RS_type.ϖ_Cabannes[iB] (1 for noRS) */
    /* host, double */
    const double*  mu;         /* [N] qp_μN */
    const double*  wt;         /* [N] wt_μN */
    const double*  aer_omega;  /* [nAer] ω̃ */
    const double*  aer_ft;     /* [nAer] fᵗ */
    const double*  tau_aer;    /* (nAer,Nz) */
    const double*  vza_weight; /* (nStokes,nVZA,n_moments): weight·(cos mφ | sin mφ)
                                  (_precompute_vza_weights, postprocessing_vza.jl:23-37) */
    /* host, int32 -- control flow derived from FULL-band reductions */
    const int32_t* ndoubl;       /* [Nz] */
    const int32_t* scatter;      /* [Nz] */
    const int32_t* iface;        /* [Nz]; the surface step reuses iface[Nz-1] (rt_run.jl:459-465) */
    const int32_t* rayl_allzero; /* [Nz] all(τ_rayl[:,iz].*ϖ_Cabannes .== 0) (types.jl:1279) */
    const int32_t* vza_stream;   /* [nVZA] 1-based nearest stream index iμ of each VZA */
    /* device, FT */
    const void* tau_rayl;    /* (nSpec,Nz) */
    const void* tau_abs;     /* (nSpec,Nz) */
    const void* F0;          /* (nStokes,nSpec) */
    const void* Zpp;         /* (N,N,1+nAer,n_moments); component 0 = Rayleigh
                                (Scattering.compute_Z_moments, compute_Z_matrices.jl:26-110) */
    const void* Zmp;
    /* device, FT -- only read when surface_kind == VSM_SURFACE_BRDF_FOURIER (and struct_size covers it):
     * (N,N,n_moments), moment m = the reference's R_surf before the Diagonal(qp_μN.*wt_μN) scaling, i.e.
     * 2*reflectance(brdf,pol_type,qp_μ,0) for m = 0 and reflectance(brdf,pol_type,qp_μ,m) otherwise
     * (rpv_surface.jl:69-75).  Spectrally constant, like the reference's. */
    const void* surf_rho;
    /* host, double, may be NULL -- only read when struct_size covers it: (nVZA,nStokes) geometry factor of the
     * Cox-Munk TMS single-scattering correction, M_exact[si,1] - M_fourier[si,1] per viewing geometry
     * (apply_ss_correction!, Surfaces/coxmunk_surface.jl:509-555, called at rt_run.jl:520-524).  The library adds
     * R_SFI[iv,si,s] += mu0 * exp(-tau_sum_all[s,end]/mu0) * ss_correction[iv,si] in its output reduction. */
    const double* ss_correction;
    /* device, FT, may be NULL -- only read when struct_size covers them: extra sources of the v0.7 source-term
     * seam (per-source j0 slots, doubling.jl:66-80, interaction.jl:231-259).
     *  thermal_B (Nz,nSpec): ThermalEmission.B_layer, the Planck radiance of every layer.  At m = 0 each scattering
     *    elemental layer gets the volume source 2 pi (1-varpi) B (1-exp(-dtau/mu_i)) on its Stokes-I rows
     *    (contribute!, Sources/thermal_emission.jl:236-316) in its own slot, doubled with expk = 1, then joined
     *    with the solar slot (the interaction step treats all slots alike and the outputs add, postprocessing_vza.jl:67-80).
     *  sif0 (nStokes,nSpec): SurfaceSIF.SIF0, injected as 2*SIF0 into the Lambertian surface's j0- at m = 0
     *    (surface_source_contribute!, Sources/surface_sif.jl:168-178). */
    const void* thermal_B;
    const void* sif0;
} vsm_column_desc;
/* outputs: device; R_SFI/T_SFI/hdr (nVZA,nStokes,nSpec); bhr_* (nStokes,nSpec);
 * hdr/bhr_uw/bhr_dw may be NULL.  Outputs are overwritten (not accumulated).  */
int vsm_rt_column_f64(vsm_ctx*, const vsm_column_desc*, double* R_SFI, double* T_SFI,
                      double* hdr, double* bhr_uw, double* bhr_dw);
int vsm_rt_column_f32(vsm_ctx*, const vsm_column_desc*, float* R_SFI, float* T_SFI,
                      float* hdr, float* bhr_uw, float* bhr_dw);
/* The same column with the rt_run_streams export (rt_run.jl:107-193, StreamRTResult): per Fourier moment
 * the composite after the surface step at ALL quadrature streams, device arrays, each may be NULL:
 *   Rmp, Tpp  (N,N,nSpec,n_moments)   composite_layer.R⁻⁺ / T⁺⁺
 *   Jm,  Jp   (N,nSpec,n_moments)     composite_layer.J₀⁻ / J₀⁺  (the reference's (N,1,nSpec) per moment) */
int vsm_rt_column_streams_f64(vsm_ctx*, const vsm_column_desc*, double* R_SFI, double* T_SFI, double* hdr,
                              double* bhr_uw, double* bhr_dw, double* Rmp, double* Tpp, double* Jm, double* Jp);
int vsm_rt_column_streams_f32(vsm_ctx*, const vsm_column_desc*, float* R_SFI, float* T_SFI, float* hdr,
                              float* bhr_uw, float* bhr_dw, float* Rmp, float* Tpp, float* Jm, float* Jp);
/* -- L2, LinMode: forward R_SFI/T_SFI and the analytic Jacobians for P parameters -- drop-in for the body of
 * rt_run(model, lin_model, ...) src/CoreRT/rt_run_lin.jl:102-326 with rt_kernel! of CoreKernel/rt_kernel_lin.jl:50-180
 * (lin elemental!, doubling_allparams!, lin interaction_helper!, lin postprocessing_vza!).  A parameter p is given
 * by the derivatives of the column inputs with respect to it (the reference's chain rule from its
 * 7*NAer + NGas + NSurf parameter vector, parameter_layout.jl:32-34, to tau-dot / varpi-dot / Z-dot per layer is
 * evaluated on the device from these): parameters that act through the layer optical depths -- Rayleigh, gas
 * absorption (per-layer VMR scalings), aerosol optical depth (tau_ref) -- and through the Lambertian albedo.
 * Parameters that change the phase matrix itself (aerosol microphysics) and BRDF shape parameters are not covered.
 * fp64 only; every layer must scatter (all interfaces _11), as in any Rayleigh atmosphere. */
typedef struct vsm_lin_desc {
    int32_t struct_size;      /* sizeof(vsm_lin_desc) */
    int32_t n_params;         /* P, 1..16 */
    const double* dtau_aer;   /* host (nAer,Nz,P) d tau_aer / dp, or NULL */
    const double* dalbedo;    /* host [P] d albedo / dp, or NULL */
    const void*   dtau_rayl;  /* device (nSpec,Nz,P) d tau_rayl / dp, or NULL */
    const void*   dtau_abs;   /* device (nSpec,Nz,P) d tau_abs / dp, or NULL */
} vsm_lin_desc;
/* outputs: device; R_SFI/T_SFI (nVZA,nStokes,nSpec); dR/dT (nVZA,nStokes,nSpec,P) as rt_run_lin.jl:185-190 */
int vsm_rt_column_lin_f64(vsm_ctx*, const vsm_column_desc*, const vsm_lin_desc*, double* R_SFI, double* T_SFI,
                          double* dR, double* dT);

/* -- L2, rotational Raman scattering (RRS): elastic + first-order inelastic radiances -- drop-in for the body of
 * rt_run(RS_type::RRS, model, iBand) src/CoreRT/rt_run.jl:238-539 with rt_kernel!(::RRS) of
 * CoreKernel/rt_kernel.jl:352-391: elemental_inelastic! (elemental_inelastic.jl:23-206, 479-603), elemental!,
 * doubling_helper!(::RRS) (doubling_inelastic.jl:13-164), interaction_helper!(::RRS, ::ScatteringInterface_11)
 * (interaction_inelastic.jl:319-521), copy_added_to_composite_ie!, postprocessing_vza!(::RRS)
 * (postprocessing_vza.jl:121-162).  The Raman line physics (getRamanSSProp!, raman_atmo_prop.jl:80-100) is the
 * caller's: it supplies the shift table i_λ₁λ₀, the normalised weights ϖ_λ₁λ₀ (normalize_raman_weights!,
 * Inelastic/types.jl:625-629) and the Z moments of greek_raman (computeRamanZλ!, inelastic_helper.jl:917-924).
 * The column descriptor describes the WHOLE band (nSpec = band length: the donor wavelength n0 = n1 + i_shift[dn]
 * of every coupling must be addressable); its Zpp/Zmp component 0 must be the Cabannes set and varpi_cabannes
 * ϖ_Cabannes (compEffectiveLayerProperties.jl:36-41).  The call computes the outputs of the OWNER range
 * n1 in [n1_begin, n1_begin + n1_count) -- the whole band for a single device, one contiguous shard per device
 * when the band is sharded: the inelastic terms read only ELASTIC state of other wavelengths, so a shard needs no
 * communication, just the elastic column of its halo [n1_begin + min shift, n1_end + max shift], which the library
 * recomputes locally.  Every layer must scatter (rt_kernel.jl:366 hard-codes scatter = true; the
 * reference has no working _01/_10 RRS interaction).  fp64, N <= 48.  thermal_B / sif0 / ss_correction must be NULL. */
typedef struct vsm_rrs_desc {
    int32_t struct_size;        /* sizeof(vsm_rrs_desc) */
    int32_t n_raman;            /* length of i_λ₁λ₀ */
    int32_t n1_begin;           /* 0-based first owned band index */
    int32_t n1_count;           /* number of owned band indices */
    const int32_t* i_shift;     /* host [n_raman]  RS_type.i_λ₁λ₀ */
    const double*  varpi_raman; /* host [n_raman]  RS_type.ϖ_λ₁λ₀ after normalize_raman_weights! */
    const void*    Zpp_raman;   /* device (N,N,n_moments)  RS_type.Z⁺⁺_λ₁λ₀ per Fourier moment */
    const void*    Zmp_raman;   /* device (N,N,n_moments)  RS_type.Z⁻⁺_λ₁λ₀ */
} vsm_rrs_desc;
/* outputs: device, (nVZA,nStokes,n1_count) each, overwritten: the elastic R_SFI/T_SFI and the inelastic
 * ieR_SFI/ieT_SFI (rt_run.jl:532) of the owner range.  vsm_last_column_ms reports the inelastic-unit kernels as
 * doubling_ms and the elastic-export kernels as adding_ms. */
int vsm_rt_column_rrs_f64(vsm_ctx*, const vsm_column_desc*, const vsm_rrs_desc*, double* R_SFI, double* T_SFI,
                          double* ieR_SFI, double* ieT_SFI);

/* -- L2, single scattering -- drop-in for the body of rt_run_ss(::noRS, model, iBand) src/CoreRT/rt_run.jl:593-755:
 * rt_kernel_ss!(::noRS{FT}) (CoreKernel/rt_kernel_ss.jl:156-236: elemental! of the whole layer with ndoubl = 0, no
 * doubling), interaction_ss! (CoreKernel/interaction_ss.jl:14-42), the surface sources, postprocessing_vza! and the
 * m = 0 hemispheric sums.  Same column descriptor (ndoubl / scatter / iface are not read: the driver hard-codes
 * scatter = true, rt_kernel_ss.jl:177); thermal_B / sif0 must be NULL.  Outputs: device, R_SFI/T_SFI
 * (nVZA,nStokes,nSpec), hem_R/hem_T [nSpec] (may be NULL), overwritten. */
int vsm_rt_column_ss_f64(vsm_ctx*, const vsm_column_desc*, double* R_SFI, double* T_SFI, double* hem_R, double* hem_T);
int vsm_rt_column_ss_f32(vsm_ctx*, const vsm_column_desc*, float* R_SFI, float* T_SFI, float* hem_R, float* hem_T);

/* The same driver with rotational Raman: rt_run_ss(RS_type::RRS, model, iBand) -- rt_kernel_ss!(::RRS)
 * (CoreKernel/rt_kernel_ss.jl:240-331: elemental_inelastic! and elemental! of the whole layer with ndoubl = 0) and
 * interaction_ss! / get_interaction_ss_RRS! (CoreKernel/interaction_ss.jl:14-19, 70-90).  The rrs descriptor as for
 * vsm_rt_column_rrs_f64 with the owner range = the whole band (n1_begin = 0, n1_count = nSpec).  Outputs: device,
 * R_SFI/T_SFI/ieR_SFI/ieT_SFI (nVZA,nStokes,nSpec), hem_R/hem_T [nSpec] (may be NULL). */
int vsm_rt_column_ss_rrs_f64(vsm_ctx*, const vsm_column_desc*, const vsm_rrs_desc*, double* R_SFI, double* T_SFI,
                             double* ieR_SFI, double* ieT_SFI, double* hem_R, double* hem_T);

/* -- several GPUs from ONE host process (the reference drives everything from a single Julia process and has no
 * multi-GPU path: Architectures.jl:42-45 "Run on a single NVIDIA CUDA GPU").  vsm_multi_create opens one context
 * (own stream) per listed device and an in-process NCCL communicator over them (ncclCommInitAll; NCCL is loaded
 * with dlopen, VSM_ERR_UNSUPPORTED if libnccl.so.2 is absent).  vsm_rt_column_sharded_* runs the column of
 * vsm_rt_column_* with the spectral axis cut into contiguous shards, one per device, no data-path collective (the
 * ndoubl / scatter / iface arrays of the descriptor must come from the FULL band, as always), ONE ncclAllGather of
 * the packed outputs, and returns the full-band result.  Differences to vsm_rt_column_*: the descriptor's
 * tau_rayl, tau_abs, F0, Zpp, Zmp and surf_rho and all five outputs are HOST pointers (the library stages the
 * shards), struct_size must be sizeof(vsm_column_desc), and the call blocks until the result is in host memory.
 * Results are bit-identical to the single-device call. */
typedef struct vsm_multi vsm_multi;
int  vsm_multi_create(vsm_multi** out, const int* devices, int ndev);
int  vsm_multi_destroy(vsm_multi*);
const char* vsm_multi_last_error(vsm_multi*);       /* may be NULL: last create error */
int  vsm_multi_ndev(vsm_multi*);
vsm_ctx* vsm_multi_ctx(vsm_multi*, int rank);       /* the per-device context (e.g. for vsm_set_inverse_mode) */
int vsm_rt_column_sharded_f64(vsm_multi*, const vsm_column_desc*, double* R_SFI, double* T_SFI, double* hdr,
                              double* bhr_uw, double* bhr_dw);
int vsm_rt_column_sharded_f32(vsm_multi*, const vsm_column_desc*, float* R_SFI, float* T_SFI, float* hdr,
                              float* bhr_uw, float* bhr_dw);

/* The Raman column (vsm_rt_column_rrs_f64) on the devices of a vsm_multi from ONE process: every device receives the whole
 * band and one contiguous owner range of the outputs (n1_begin / n1_count of the descriptor are ignored), recomputes the
 * elastic column of its halo locally (no data-path collective) and ONE ncclAllGather of the packed (R | T | ieR | ieT)
 * rows closes the call.  tau_rayl, tau_abs, F0, Zpp, Zmp, surf_rho, Zpp_raman, Zmp_raman and the four outputs
 * (nVZA,nStokes,nSpec) are HOST pointers; blocks until the result is in host memory; bit-identical to one device. */
int vsm_rt_column_rrs_sharded_f64(vsm_multi*, const vsm_column_desc*, const vsm_rrs_desc*, double* R_SFI, double* T_SFI,
                                  double* ieR_SFI, double* ieT_SFI);

/* device time (ms) of the most recent vsm_rt_column_* call, measured with CUDA
 * events on the context's stream; blocks until that call has finished.        */
int vsm_last_column_ms(vsm_ctx*, float* total_ms, float* doubling_ms, float* adding_ms);
/* failure detection for the fused column (the reference allocates cuBLAS' info arrays but never reads them,
 * gpu_batched_cuda.jl:146-153): the number of R_SFI / T_SFI elements of the most recent vsm_rt_column_* call that
 * are not finite -- a singular (I - A) or an overflow anywhere in a spectral point's chain ends up there.  Blocks
 * until that call has finished.  (vsm_batch_inv_* reports per-slice zero pivots through its info array.) */
int vsm_last_column_status(vsm_ctx*, int* nonfinite_outputs);
/* number of k_layer (elemental+doubling) and k_adding launches of that call */
int vsm_last_column_launches(vsm_ctx*, int* layer_launches, int* adding_launches);

#ifdef __cplusplus
}
#endif
#endif /* VSMOM_B200_H */

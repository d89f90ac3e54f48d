# vSmartMOMB200Ext.jl -- package-extension glue a vSmartMOM maintainer would add under ext/
# (Project.toml: [extensions] vSmartMOMB200Ext = "CUDA").  It installs, for CuArray inputs,
# the same methods ext/gpu_batched_cuda.jl installs (L0), a more specific rt_kernel! (L1)
# and an rt_run body replacement (L2), all forwarding to libvsmom_b200.so.
# NOT executable in the build container (no Julia); mirrored 1:1 by the Python package
# vsmartmom.jl_b200/vsmom_b200 which the tests drive through the same C-ABI.
module vSmartMOMB200Ext

using CUDA
using vSmartMOM
using vSmartMOM.CoreRT
using vSmartMOM.Architectures

const LIB = get(ENV, "VSMOM_B200_LIB", "libvsmom_b200")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

_check(rc) = rc == 0 || error("libvsmom_b200: " *
    unsafe_string(ccall((:vsm_last_error, LIB), Cstring, (Ptr{Cvoid},), CTX[])))

function __init__()                                         # ext/vSmartMOMCUDAExt.jl:36-66
    CUDA.functional() || return
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:vsm_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}),
               h, CUDA.deviceid(), CUDA.stream().handle)
    rc == 0 || (@warn "vsmom_b200 unavailable, keeping the cuBLAS path"; return)
    CTX[] = h[]
    Architectures._has_cuda[] = true
    Architectures._sync_gpu[] = () -> ccall((:vsm_synchronize, LIB), Cint, (Ptr{Cvoid},), CTX[])
end

sfx(::Type{Float64}) = "f64"; sfx(::Type{Float32}) = "f32"

# ---- L0: ext/gpu_batched_cuda.jl:208-218 / :97-182 / :72-94 -----------------------------
for FT in (Float64, Float32)
    mul, inv_, sol = Symbol("vsm_batched_mul_", sfx(FT)), Symbol("vsm_batch_inv_", sfx(FT)), Symbol("vsm_batch_solve_", sfx(FT))
    @eval function CoreRT.batched_mul(A::CuArray{$FT,3}, B::CuArray{$FT,3})
        C = CuArray{$FT,3}(undef, size(A,1), size(B,2), size(A,3))      # fresh result, like the reference
        _check(ccall(($(QuoteNode(mul)), LIB), Cint,
            (Ptr{Cvoid}, CuPtr{$FT}, CuPtr{$FT}, CuPtr{$FT}, Cint, Cint, Cint, Cint),
            CTX[], C, A, B, size(A,1), size(A,2), size(B,2), size(A,3)))
        return C
    end
    @eval function CoreRT.batch_inv!(X::CuArray{$FT,3}, A::CuArray{$FT,3}, args...)   # all 3 arities
        _check(ccall(($(QuoteNode(inv_)), LIB), Cint,
            (Ptr{Cvoid}, CuPtr{$FT}, CuPtr{$FT}, Cint, Cint, CuPtr{Cint}),
            CTX[], X, A, size(A,1), size(A,3), CU_NULL))
        return X
    end
    @eval function CoreRT.batch_solve!(X::CuArray{$FT,3}, A::CuArray{$FT,3}, B::CuArray{$FT,3})
        _check(ccall(($(QuoteNode(sol)), LIB), Cint,
            (Ptr{Cvoid}, CuPtr{$FT}, CuPtr{$FT}, CuPtr{$FT}, Cint, Cint, Cint),
            CTX[], X, A, B, size(A,1), size(B,2), size(A,3)))
    end
end
CoreRT.batched_pointer_cache(::CuArray) = nothing            # pointer arrays are not needed

# ---- L1: more specific method of rt_kernel!(::noRS, ...) (rt_kernel.jl:175-189) ---------
Base.@kwdef mutable struct LayerDesc                          # == struct vsm_layer_desc
    struct_size::Int32 = 0; N::Int32 = 0; nStokes::Int32 = 0; nSpec::Int32 = 0; m::Int32 = 0
    imu0::Int32 = 0; ndoubl::Int32 = 0; scatter::Int32 = 0; iface::Int32 = 0; iz::Int32 = 0
    z_is_spectral::Int32 = 0; reserved::Int32 = 0; mu0::Float64 = 0
    ptrs::NTuple{20,CuPtr{Cvoid}} = ntuple(_ -> CU_NULL, 20)  # tau … j0m in header order
end
iface_tag(::CoreRT.ScatteringInterface_00) = 0; iface_tag(::CoreRT.ScatteringInterface_01) = 1
iface_tag(::CoreRT.ScatteringInterface_10) = 2; iface_tag(::CoreRT.ScatteringInterface_11) = 3

function CoreRT.rt_kernel!(RS::CoreRT.noRS{FT}, pol_type, SFI, added, comp,
        props::CoreRT.CoreScatteringOpticalProperties, interface, τ_sum, m, qp, I_static,
        ::Architectures.GPU, qp_μN, iz; dτ_max_threshold=nothing, dτ_min_floor=nothing, kw...) where {FT}
    scatter = maximum(props.τ .* props.ϖ) > 2eps(FT)                          # rt_kernel.jl:197
    _, nd = scatter ? CoreRT.get_dtau_ndoubl(props, qp; dτ_max_threshold, dτ_min_floor) : (nothing, 0)
    p(x) = reinterpret(CuPtr{Cvoid}, pointer(x))
    d = LayerDesc(struct_size = sizeof(LayerDesc), N = length(qp_μN), nStokes = pol_type.n,
        nSpec = length(props.τ), m = m, imu0 = qp.iμ₀, ndoubl = nd, scatter = scatter,
        iface = iface_tag(interface), iz = iz, z_is_spectral = size(props.Z⁺⁺, 3) > 1, mu0 = qp.μ₀,
        ptrs = (p(props.τ), p(props.ϖ), p(props.Z⁺⁺), p(props.Z⁻⁺), p(CuArray(τ_sum)), p(CuArray(RS.F₀)),
                p(qp.qp_μN), p(qp.wt_μN), p(comp.R⁻⁺), p(comp.R⁺⁻), p(comp.T⁺⁺), p(comp.T⁻⁻),
                p(comp.J₀⁺), p(comp.J₀⁻), p(added.r⁻⁺), p(added.r⁺⁻), p(added.t⁺⁺), p(added.t⁻⁻),
                p(added.j₀⁺), p(added.j₀⁻)))
    _check(ccall((Symbol("vsm_rt_layer_", sfx(FT)), LIB), Cint, (Ptr{Cvoid}, Ref{LayerDesc}), CTX[], d))
end

# ---- L2: vsm_rt_column_* replaces the m / iz loops of rt_run(::noRS, model, iBand)
# (rt_run.jl:383-517).  Surfaces: LambertianSurfaceScalar natively (surface_kind 0), every other
# AbstractSurfaceType through the Fourier moments of R_surf that create_surface_layer! builds on
# the host anyway (surface_kind 1, surf_rho = (m == 0 ? 2 : 1) * reflectance(brdf, pol_type, qp_μ, m),
# rpv_surface.jl:69-75).  vsm_rt_column_streams_* additionally fills the StreamRTResult arrays of
# rt_run_streams (rt_run.jl:107-193).  vsm_set_inverse_mode(ctx, 0) forces the elimination route
# for every (I - A)^-1 (default 1: adaptive Neumann / Gauss-Jordan).  The descriptor is filled from
# model.τ_rayl/τ_abs/τ_aer, Scattering.compute_Z_moments for m = 0:m_max, the full-band
# get_dtau_ndoubl / extractEffectiveProps control flow and _precompute_vza_weights --
# exactly what vsmom_b200/rt_run.py::rt_column_device does; see INTEGRATION.md §3.
end # module

# vSmartMOMB200Ext.jl -- package extension that REPLACES ext/vSmartMOMCUDAExt.jl.
#
#   Project.toml          [extensions]
#                         vSmartMOMB200Ext = "CUDA"          # instead of: vSmartMOMCUDAExt = "CUDA"
#
# Both extensions define the same CuArray methods of CoreRT.batched_mul / batch_inv! / ... , and two
# extensions triggered by the same weak dependency cannot both define a method (method overwriting is an
# error during precompilation).  This file therefore provides EVERYTHING vSmartMOMCUDAExt provides
# (ext/vSmartMOMCUDAExt.jl:21-66, ext/gpu_batched_cuda.jl:65-271), so that every path of vSmartMOM keeps
# working on the GPU, and on top of it the three levels of include/vsmom_b200.h:
#   L0  batched_mul / batch_inv! / batch_solve!        -> vsm_batched_mul / vsm_batch_inv / vsm_batch_solve
#   L1  rt_kernel!(::noRS, ..., ::Architectures.GPU)   -> vsm_rt_layer            (rt_kernel.jl:175-250)
#   L2  rt_run(::noRS, ::RTModel{<:GPU}, iBand)        -> vsm_rt_column[_streams] (rt_run.jl:238-539)
# SubArray and ForwardDiff.Dual inputs (gpu_batched_cuda.jl:221-271) are reduced to the dense Float methods.
#
# Julia is not installed in the build container: this file has not been executed.  The Python package
# vsmartmom.jl_b200/vsmom_b200 drives the same entry points with the same argument meaning through ctypes
# and is what the tests exercise; rt_run.py::_column_desc is the executable form of `column_desc` below.
module vSmartMOMB200Ext

using CUDA
using ForwardDiff
using LinearAlgebra
using vSmartMOM
using vSmartMOM.Architectures
using vSmartMOM.CoreRT
using vSmartMOM.Scattering
using vSmartMOM.InelasticScattering
import CUDA.CUDAKernels

const LIB = get(ENV, "VSMOM_B200_LIB", "libvsmom_b200")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
"Read the per-slice singular flags of batch_inv! back (one small D2H per call; the reference never looks at them)."
const CHECK_INFO = Ref(true)

function _check(rc)
    rc == 0 && return nothing
    error("libvsmom_b200 (code $rc): " * unsafe_string(ccall((:vsm_last_error, LIB), Cstring, (Ptr{Cvoid},), CTX[])))
end

# ---- what vSmartMOMCUDAExt installs (ext/vSmartMOMCUDAExt.jl:21-27) ------------------------------------
Architectures.devi(::Architectures.GPU) = CUDA.CUDABackend(; always_inline=true)
Architectures.array_type(::Architectures.GPU) = CuArray
Architectures.architecture(::CuArray) = Architectures.GPU()

function __init__()                                          # ext/vSmartMOMCUDAExt.jl:36-66
    Architectures._has_cuda[] = false
    CUDA.functional() || return
    if Base.get_extension(vSmartMOM, :vSmartMOMCUDAExt) !== nothing
        error("vSmartMOMB200Ext replaces vSmartMOMCUDAExt: list only one of them under [extensions]")
    end
    try
        CoreRT.CUBLAS_ref[] = CUDA.CUBLAS                    # batched_pointer_cache, Dual / fallback paths
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:vsm_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}),
                   h, CUDA.deviceid(), CUDA.stream().handle)
        rc == 0 || error(unsafe_string(ccall((:vsm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        CTX[] = h[]
        Architectures._has_cuda[] = true
        Architectures._sync_gpu[] = CUDA.synchronize         # library work is on CUDA.jl's task stream
        CUDA.allowscalar(false)
        atexit(() -> (CTX[] == C_NULL || ccall((:vsm_destroy, LIB), Cint, (Ptr{Cvoid},), CTX[]); CTX[] = C_NULL))
    catch e
        @warn "vSmartMOM B200 initialization failed, falling back to CPU" exception = e
        Architectures._has_cuda[] = false
    end
end

sfx(::Type{Float64}) = "f64"
sfx(::Type{Float32}) = "f32"
const FTs = (Float64, Float32)

# ---- L0 (ext/gpu_batched_cuda.jl:65-233) ----------------------------------------------------------------
CoreRT.batched_pointer_cache(A::CuArray) = CUDA.CUBLAS.unsafe_strided_batch(A)      # :65-69 (kept: AddedLayer stores it)

for FT in FTs
    mul = Symbol("vsm_batched_mul_", sfx(FT))
    inv_ = Symbol("vsm_batch_inv_", sfx(FT))
    sol = Symbol("vsm_batch_solve_", sfx(FT))
    @eval function CoreRT.batched_mul(A::CuArray{$FT,3}, B::CuArray{$FT,3})          # :208-218 (batch == 1 included)
        C = CuArray{$FT,3}(undef, size(A, 1), size(B, 2), size(A, 3))                # fresh result, like the reference
        GC.@preserve A B C _check(ccall(($(QuoteNode(mul)), LIB), Cint,
            (Ptr{Cvoid}, CuPtr{$FT}, CuPtr{$FT}, CuPtr{$FT}, Cint, Cint, Cint, Cint),
            CTX[], C, A, B, size(A, 1), size(A, 2), size(B, 2), size(A, 3)))
        return C
    end
    @eval function _b200_inv!(X::CuArray{$FT,3}, A::CuArray{$FT,3})
        info = CHECK_INFO[] ? CUDA.zeros(Cint, size(A, 3)) : nothing
        GC.@preserve X A info _check(ccall(($(QuoteNode(inv_)), LIB), Cint,
            (Ptr{Cvoid}, CuPtr{$FT}, CuPtr{$FT}, Cint, Cint, CuPtr{Cint}),
            CTX[], X, A, size(A, 1), size(A, 3), info === nothing ? CU_NULL : pointer(info)))
        if info !== nothing
            bad = findfirst(!=(0), Array(info))
            bad === nothing || throw(LinearAlgebra.SingularException(bad))
        end
        return X
    end
    # the three arities of the reference (:97-130 plain / with RTWorkspace, :133-182 with pointer caches)
    @eval CoreRT.batch_inv!(X::CuArray{$FT,3}, A::CuArray{$FT,3}) = _b200_inv!(X, A)
    @eval CoreRT.batch_inv!(X::CuArray{$FT,3}, A::CuArray{$FT,3}, ws::CoreRT.RTWorkspace) = _b200_inv!(X, A)
    @eval CoreRT.batch_inv!(X::CuArray{$FT,3}, A::CuArray{$FT,3}, Xptrs, Aptrs) = _b200_inv!(X, A)
    @eval function CoreRT.batch_solve!(X::CuArray{$FT,3}, A::CuArray{$FT,3}, B::CuArray{$FT,3})   # :72-94
        GC.@preserve X A B _check(ccall(($(QuoteNode(sol)), LIB), Cint,
            (Ptr{Cvoid}, CuPtr{$FT}, CuPtr{$FT}, CuPtr{$FT}, Cint, Cint, Cint),
            CTX[], X, A, B, size(A, 1), size(B, 2), size(A, 3)))
        return X
    end
end

# views are materialised first (:221-233)
const CuView3{FT} = SubArray{FT,3,<:CuArray}
_dense(A::CuArray) = A
_dense(A::CuView3) = copy(A)
CoreRT.batched_mul(A::CuView3{FT}, B::CuArray{FT,3}) where {FT} = CoreRT.batched_mul(_dense(A), B)
CoreRT.batched_mul(A::CuArray{FT,3}, B::CuView3{FT}) where {FT} = CoreRT.batched_mul(A, _dense(B))
CoreRT.batched_mul(A::CuView3{FT}, B::CuView3{FT}) where {FT} = CoreRT.batched_mul(_dense(A), _dense(B))

# ForwardDiff.Dual (:236-271): value and partials go through the Float methods above
const CuDual3{T,V,N} = CuArray{ForwardDiff.Dual{T,V,N},3}
function CoreRT.batched_mul(A::CuDual3{T,V,N}, B::CuDual3{T,V,N}) where {T,V,N}
    Av, Bv = ForwardDiff.value.(A), ForwardDiff.value.(B)
    Cv = CoreRT.batched_mul(Av, Bv)
    dC = ntuple(i -> CoreRT.batched_mul(Av, ForwardDiff.partials.(B, i)) .+
                     CoreRT.batched_mul(ForwardDiff.partials.(A, i), Bv), N)
    return eltype(A).(Cv, ForwardDiff.Partials.(tuple.(dC...)))
end
function CoreRT.batch_inv!(X::CuDual3{T,V,N}, A::CuDual3{T,V,N}, args...) where {T,V,N}
    Av = ForwardDiff.value.(A)
    Ai = similar(Av)
    CoreRT.batch_inv!(Ai, copy(Av))
    dX = ntuple(i -> .-CoreRT.batched_mul(CoreRT.batched_mul(Ai, ForwardDiff.partials.(A, i)), Ai), N)   # -A^-1 A' A^-1
    X .= eltype(X).(Ai, ForwardDiff.Partials.(tuple.(dX...)))
    return X
end

function CoreRT.make_gpu_rt_workspace(FT::Type, NquadN::Int, nSpec::Int)           # :189-205
    z3() = CUDA.zeros(FT, NquadN, NquadN, nSpec)
    zJ() = CUDA.zeros(FT, NquadN, 1, nSpec)
    CoreRT.RTWorkspace(z3(), z3(), zJ(), zJ(), CUDA.zeros(Cint, NquadN, nSpec), CUDA.zeros(Cint, nSpec),
                       z3(), z3(), z3(), z3())
end

# ---- L1: rt_kernel!(::noRS, ...) on the GPU (rt_kernel.jl:175-250) ---------------------------------------
Base.@kwdef mutable struct LayerDesc                          # == struct vsm_layer_desc
    struct_size::Int32 = 0; N::Int32 = 0; nStokes::Int32 = 0; nSpec::Int32 = 0; m::Int32 = 0
    imu0::Int32 = 0; ndoubl::Int32 = 0; scatter::Int32 = 0; iface::Int32 = 0; iz::Int32 = 0
    z_is_spectral::Int32 = 0; reserved::Int32 = 0; mu0::Float64 = 0
    ptrs::NTuple{20,CuPtr{Cvoid}} = ntuple(_ -> CU_NULL, 20)  # tau ... j0m in header order
end
iface_tag(::CoreRT.ScatteringInterface_00) = 0
iface_tag(::CoreRT.ScatteringInterface_01) = 1
iface_tag(::CoreRT.ScatteringInterface_10) = 2
iface_tag(::CoreRT.ScatteringInterface_11) = 3

# device copies of small host arrays that rt_run hands to every layer call (F0, quadrature): uploaded once
# per distinct host array, not once per layer and moment
const _DEVCACHE = IdDict{Any,Any}()
_dev(x::CuArray) = x
_dev(x::AbstractArray) = get!(() -> CuArray(x), _DEVCACHE, x)
_p(x) = reinterpret(CuPtr{Cvoid}, pointer(x))

function CoreRT.rt_kernel!(RS::CoreRT.noRS{FT}, pol_type, SFI, added, comp,
        props::CoreRT.CoreScatteringOpticalProperties, interface, τ_sum, m, qp, I_static,
        ::Architectures.GPU, qp_μN, iz; dτ_max_threshold=nothing, dτ_min_floor=nothing, kw...) where {FT<:Union{Float32,Float64}}
    scatter = maximum(props.τ .* props.ϖ) > 2eps(FT)                          # rt_kernel.jl:197
    nd = scatter ? CoreRT.get_dtau_ndoubl(props, qp; dτ_max_threshold, dτ_min_floor)[2] : 0
    τs = CuArray{FT}(τ_sum)                                                   # changes per layer: a fresh upload, kept alive below
    F0, μd, wd = _dev(RS.F₀), _dev(qp.qp_μN), _dev(qp.wt_μN)
    d = LayerDesc(struct_size=sizeof(LayerDesc), N=length(qp_μN), nStokes=pol_type.n,
        nSpec=length(props.τ), m=m, imu0=qp.iμ₀, ndoubl=nd, scatter=scatter,
        iface=iface_tag(interface), iz=iz, z_is_spectral=size(props.Z⁺⁺, 3) > 1, mu0=qp.μ₀,
        ptrs=(_p(props.τ), _p(props.ϖ), _p(props.Z⁺⁺), _p(props.Z⁻⁺), _p(τs), _p(F0), _p(μd), _p(wd),
              _p(comp.R⁻⁺), _p(comp.R⁺⁻), _p(comp.T⁺⁺), _p(comp.T⁻⁻), _p(comp.J₀⁺), _p(comp.J₀⁻),
              _p(added.r⁻⁺), _p(added.r⁺⁻), _p(added.t⁺⁺), _p(added.t⁻⁻), _p(added.j₀⁺), _p(added.j₀⁻)))
    GC.@preserve props τs F0 μd wd comp added d begin
        _check(ccall((FT === Float64 ? :vsm_rt_layer_f64 : :vsm_rt_layer_f32, LIB), Cint, (Ptr{Cvoid}, Ref{LayerDesc}), CTX[], d))
        Architectures.synchronize_if_gpu()                                    # τs may be freed after this point
    end
    return nothing
end

# ---- L2: the whole band in one call (loop body of rt_run, rt_run.jl:383-524) -----------------------------
mutable struct ColumnDesc                                     # == struct vsm_column_desc
    struct_size::Int32; N::Int32; nStokes::Int32; nSpec::Int32; Nz::Int32; nAer::Int32; nVZA::Int32
    n_moments::Int32; imu0::Int32; surface_kind::Int32
    mu0::Float64; albedo::Float64; varpi_cabannes::Float64
    mu::Ptr{Float64}; wt::Ptr{Float64}; aer_omega::Ptr{Float64}; aer_ft::Ptr{Float64}; tau_aer::Ptr{Float64}
    vza_weight::Ptr{Float64}
    ndoubl::Ptr{Int32}; scatter::Ptr{Int32}; iface::Ptr{Int32}; rayl_allzero::Ptr{Int32}; vza_stream::Ptr{Int32}
    tau_rayl::CuPtr{Cvoid}; tau_abs::CuPtr{Cvoid}; F0::CuPtr{Cvoid}; Zpp::CuPtr{Cvoid}; Zmp::CuPtr{Cvoid}
    surf_rho::CuPtr{Cvoid}
    ss_correction::Ptr{Float64}
end

"True when the scene is one the library evaluates natively; anything else runs the reference's own loop on L0/L1."
function _native(RS_type, model, iBand, sources)
    length(iBand) == 1 || return false
    sources === nothing || sources isa CoreRT.SolarBeam || return false
    model.sources isa CoreRT.SolarBeam || return false
    brdf = CoreRT.get_surface(model, iBand[1])
    brdf isa CoreRT.CanopySurface && return false            # canopy sub-layers: reference path
    return true
end

function CoreRT.rt_run(RS_type::CoreRT.noRS{FT}, model::CoreRT.RTModel{<:Architectures.GPU,FT}, iBand;
                       sources=nothing, streams_callback=nothing) where {FT<:Union{Float32,Float64}}
    if CTX[] == C_NULL || !_native(RS_type, model, iBand, sources)
        return invoke(CoreRT.rt_run, Tuple{CoreRT.AbstractRamanType,Any,Any}, RS_type, model, iBand;
                      sources=sources, streams_callback=streams_callback)
    end
    iB = iBand[1]
    (; sza, vza, vaz) = model.obs_geom
    qp = model.quad_points
    (; qp_μ, qp_μN, wt_μN, μ₀, iμ₀) = qp
    pol = CoreRT.polarization_type(model)
    n, N, nV = pol.n, length(qp_μN), length(vza)
    m_max = CoreRT.m_max_bands(model)[iB]
    nM = m_max + 1
    τ_rayl, τ_abs, τ_aer = model.τ_rayl[iB], model.τ_abs[iB], model.τ_aer[iB]
    nSpec, Nz, nAer = size(τ_abs, 1), size(τ_abs, 2), size(τ_aer, 1)
    brdf = CoreRT.get_surface(model, iB)
    μ = collect(qp_μ)
    # Z-moment tables (N,N,1+nAer,nM): constructCoreOpticalProperties, compEffectiveLayerProperties.jl:36-53
    Zpp = Array{FT}(undef, N, N, 1 + nAer, nM); Zmp = similar(Zpp)
    gr = model.greek_rayleigh isa AbstractVector ? model.greek_rayleigh[iB] : model.greek_rayleigh
    for m in 0:m_max
        Zpp[:, :, 1, m+1], Zmp[:, :, 1, m+1] = Array.(Scattering.compute_Z_moments(pol, μ, gr, m))
        for a in 1:nAer
            Zpp[:, :, 1+a, m+1], Zmp[:, :, 1+a, m+1] =
                Array.(Scattering.compute_Z_moments(pol, μ, model.aerosol_optics[iB][a].greek_coefs, m))
        end
    end
    # full-band control flow (m = 0 properties: τ, ϖ do not depend on m): rt_kernel.jl:197,266-287,
    # extractEffectiveProps (compEffectiveLayerProperties.jl:75-93)
    props, _ = CoreRT.constructCoreOpticalProperties(RS_type, iBand, 0, model)
    ifaces, _ = CoreRT.extractEffectiveProps(props, qp)
    scatter = Int32[maximum(p.τ .* p.ϖ) > 2eps(FT) for p in props]
    ndoubl = Int32[scatter[i] == 1 ? CoreRT.get_dtau_ndoubl(props[i], qp;
                   dτ_max_threshold=model.numerics.dτ_max_threshold, dτ_min_floor=model.numerics.dτ_min_floor)[2] : 0
                   for i in 1:Nz]
    iface = Int32[iface_tag(x) for x in ifaces]
    rz = Int32[all(τ_rayl[:, i] .* RS_type.ϖ_Cabannes[iB] .== 0) for i in 1:Nz]      # types.jl:1279
    vstream = Int32[CoreRT.nearest_point(qp_μ, cosd(v)) for v in vza]
    vw = Array{Float64}(undef, n, nV, nM)                                           # postprocessing_vza.jl:23-37
    for m in 0:m_max, i in 1:nV
        w = m == 0 ? 0.5 / π : 1.0 / π
        c, s = cosd(m * vaz[i]), sind(m * vaz[i])
        vw[:, i, m+1] .= w .* (c, c, s, s)[1:n]
    end
    # surface: Lambertian scalar natively, every other AbstractSurfaceType through the Fourier moments of
    # R_surf exactly as create_surface_layer! forms them (rpv_surface.jl:69-75); Cox-Munk adds the TMS factor
    lamb = brdf isa CoreRT.LambertianSurfaceScalar
    ρ = lamb ? nothing : CuArray{FT}(cat([(m == 0 ? 2 : 1) .* CoreRT.reflectance(brdf, pol, μ, m) for m in 0:m_max]...; dims=3))
    ss = nothing
    if brdf isa CoreRT.CoxMunkSurface                                                # apply_ss_correction!, coxmunk_surface.jl:509-555
        ss = Array{Float64}(undef, nV, n)
        for iv in 1:nV, si in 1:n
            μv, dϕ = FT(cosd(vza[iv])), FT(deg2rad(vaz[iv]))
            four = zero(FT)
            for m in 0:m_max
                four += (m == 0 ? FT(0.5) : FT(1)) * CoreRT._azimuthal_kernel(si, 1, m, dϕ) *
                        CoreRT._fourier_coeff_element(brdf, n, si, 1, μv, μ₀, m)
            end
            ss[iv, si] = CoreRT.coxmunk_brdf_mueller(brdf, n, μv, μ₀, dϕ)[si, 1] - four
        end
    end
    ω = Float64[model.aerosol_optics[iB][a].ω̃ for a in 1:nAer]; isempty(ω) && push!(ω, 0.0)
    ft = Float64[model.aerosol_optics[iB][a].fᵗ for a in 1:nAer]; isempty(ft) && push!(ft, 0.0)
    τa = nAer > 0 ? Array{Float64}(τ_aer) : zeros(1, 1)
    μh, wh = Array{Float64}(qp_μN), Array{Float64}(wt_μN)
    F0 = size(RS_type.F₀) == (n, nSpec) ? RS_type.F₀ : (f = zeros(FT, n, nSpec); f[1, :] .= 1; f)
    d_τr, d_τa, d_F0, d_Zpp, d_Zmp = CuArray{FT}(τ_rayl), CuArray{FT}(τ_abs), CuArray{FT}(F0), CuArray(Zpp), CuArray(Zmp)
    R_d = CuArray{FT}(undef, nV, n, nSpec); T_d = similar(R_d); H_d = similar(R_d)
    uw_d = CuArray{FT}(undef, n, nSpec); dw_d = similar(uw_d)
    want_streams = streams_callback !== nothing
    sR = want_streams ? CuArray{FT}(undef, N, N, nSpec, nM) : nothing; sT = want_streams ? similar(sR) : nothing
    sJm = want_streams ? CuArray{FT}(undef, N, nSpec, nM) : nothing; sJp = want_streams ? similar(sJm) : nothing
    GC.@preserve μh wh ω ft τa vw ndoubl scatter iface rz vstream ss ρ d_τr d_τa d_F0 d_Zpp d_Zmp begin
        d = ColumnDesc(sizeof(ColumnDesc), N, n, nSpec, Nz, nAer, nV, nM, iμ₀, lamb ? 0 : 1,
            μ₀, lamb ? brdf.albedo : 0.0, RS_type.ϖ_Cabannes[iB],
            pointer(μh), pointer(wh), pointer(ω), pointer(ft), pointer(τa), pointer(vw),
            pointer(ndoubl), pointer(scatter), pointer(iface), pointer(rz), pointer(vstream),
            _p(d_τr), _p(d_τa), _p(d_F0), _p(d_Zpp), _p(d_Zmp), lamb ? CU_NULL : _p(ρ),
            ss === nothing ? Ptr{Float64}(C_NULL) : pointer(ss))
        if want_streams
            _check(ccall((FT === Float64 ? :vsm_rt_column_streams_f64 : :vsm_rt_column_streams_f32, LIB), Cint,
                (Ptr{Cvoid}, Ref{ColumnDesc}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}),
                CTX[], d, R_d, T_d, H_d, uw_d, dw_d, sR, sT, sJm, sJp))
        else
            _check(ccall((FT === Float64 ? :vsm_rt_column_f64 : :vsm_rt_column_f32, LIB), Cint,
                (Ptr{Cvoid}, Ref{ColumnDesc}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}),
                CTX[], d, R_d, T_d, H_d, uw_d, dw_d))
        end
        Architectures.synchronize_if_gpu()
    end
    if want_streams                                            # streams_callback(m, composite_layer-like views), rt_run.jl:467-480
        for m in 0:m_max
            streams_callback(m, (R⁻⁺=view(sR, :, :, :, m + 1), T⁺⁺=view(sT, :, :, :, m + 1),
                                 J₀⁻=reshape(view(sJm, :, :, m + 1), N, 1, nSpec), J₀⁺=reshape(view(sJp, :, :, m + 1), N, 1, nSpec)))
        end
    end
    R_SFI, T_SFI, hdr = Array(R_d), Array(T_d), Array(H_d)     # ONE D2H per output instead of one per Fourier moment
    z = zeros(FT, nV, n, nSpec)
    return R_SFI, T_SFI, z, copy(z), hdr, Array(uw_d)[1, :], Array(dw_d)[1, :]     # (R_SFI,T_SFI,ieR_SFI,ieT_SFI,hdr,bhr_uw,bhr_dw), rt_run.jl:535
end

# ---- L2, rotational Raman: rt_run(RS_type::RRS, model, iBand) (rt_run.jl:238-539, rt_kernel!(::RRS) rt_kernel.jl:352-391)
mutable struct RrsDesc                                        # == struct vsm_rrs_desc
    struct_size::Int32; n_raman::Int32; n1_begin::Int32; n1_count::Int32
    i_shift::Ptr{Int32}; varpi_raman::Ptr{Float64}; Zpp_raman::CuPtr{Cvoid}; Zmp_raman::CuPtr{Cvoid}
end

function CoreRT.rt_run(RS_type::InelasticScattering.RRS{Float64}, model::CoreRT.RTModel{<:Architectures.GPU,Float64}, iBand;
                       sources=nothing, streams_callback=nothing)
    FT = Float64
    brdf0 = CoreRT.get_surface(model, iBand[1])
    if CTX[] == C_NULL || !_native(RS_type, model, iBand, sources) || streams_callback !== nothing || brdf0 isa CoreRT.CoxMunkSurface
        return invoke(CoreRT.rt_run, Tuple{CoreRT.AbstractRamanType,Any,Any}, RS_type, model, iBand;
                      sources=sources, streams_callback=streams_callback)
    end
    iB = iBand[1]
    InelasticScattering.normalize_raman_weights!(RS_type, model, iBand)          # rt_run.jl:275
    (; vza, vaz) = model.obs_geom
    qp = model.quad_points
    (; qp_μ, qp_μN, wt_μN, μ₀, iμ₀) = qp
    pol = CoreRT.polarization_type(model)
    n, N, nV = pol.n, length(qp_μN), length(vza)
    m_max = CoreRT.m_max_bands(model)[iB]
    nM = m_max + 1
    τ_rayl, τ_abs, τ_aer = model.τ_rayl[iB], model.τ_abs[iB], model.τ_aer[iB]
    nSpec, Nz, nAer = size(τ_abs, 1), size(τ_abs, 2), size(τ_aer, 1)
    μ = collect(qp_μ)
    # elastic Z tables with the CABANNES Greek set (compEffectiveLayerProperties.jl:36-41) and the Raman Z moments
    # (computeRamanZλ!, inelastic_helper.jl:917-924), all Fourier moments at once
    Zpp = Array{FT}(undef, N, N, 1 + nAer, nM); Zmp = similar(Zpp)
    ZppR = Array{FT}(undef, N, N, nM); ZmpR = similar(ZppR)
    grs = CoreRT._rayleigh_greek_source(RS_type, model.greek_rayleigh, model.greek_cabannes)
    gr = grs isa AbstractVector ? grs[iB] : grs
    for m in 0:m_max
        Zpp[:, :, 1, m+1], Zmp[:, :, 1, m+1] = Array.(Scattering.compute_Z_moments(pol, μ, gr, m))
        for a in 1:nAer
            Zpp[:, :, 1+a, m+1], Zmp[:, :, 1+a, m+1] =
                Array.(Scattering.compute_Z_moments(pol, μ, model.aerosol_optics[iB][a].greek_coefs, m))
        end
        ZppR[:, :, m+1], ZmpR[:, :, m+1] = Array.(Scattering.compute_Z_moments(pol, μ, RS_type.greek_raman, m))
    end
    props, _ = CoreRT.constructCoreOpticalProperties(RS_type, iBand, 0, model)
    ifaces, _ = CoreRT.extractEffectiveProps(props, qp)
    scatter = Int32[maximum(p.τ .* p.ϖ) > 2eps(FT) for p in props]
    all(scatter .== 1) || error("vSmartMOMB200Ext: RRS needs every layer to scatter (rt_kernel.jl:366)")
    ndoubl = Int32[CoreRT.get_dtau_ndoubl(props[i], qp; dτ_max_threshold=model.numerics.dτ_max_threshold,
                                          dτ_min_floor=model.numerics.dτ_min_floor)[2] for i in 1:Nz]
    iface = Int32[iface_tag(x) for x in ifaces]
    rz = Int32[all(τ_rayl[:, i] .* RS_type.ϖ_Cabannes[iB] .== 0) for i in 1:Nz]
    vstream = Int32[CoreRT.nearest_point(qp_μ, cosd(v)) for v in vza]
    vw = Array{Float64}(undef, n, nV, nM)
    for m in 0:m_max, i in 1:nV
        w = m == 0 ? 0.5 / π : 1.0 / π
        c, s = cosd(m * vaz[i]), sind(m * vaz[i])
        vw[:, i, m+1] .= w .* (c, c, s, s)[1:n]
    end
    lamb = brdf0 isa CoreRT.LambertianSurfaceScalar
    ρ = lamb ? nothing : CuArray{FT}(cat([(m == 0 ? 2 : 1) .* CoreRT.reflectance(brdf0, pol, μ, m) for m in 0:m_max]...; dims=3))
    ω = Float64[model.aerosol_optics[iB][a].ω̃ for a in 1:nAer]; isempty(ω) && push!(ω, 0.0)
    ft = Float64[model.aerosol_optics[iB][a].fᵗ for a in 1:nAer]; isempty(ft) && push!(ft, 0.0)
    τa = nAer > 0 ? Array{Float64}(τ_aer) : zeros(1, 1)
    μh, wh = Array{Float64}(qp_μN), Array{Float64}(wt_μN)
    F0 = size(RS_type.F₀) == (n, nSpec) ? RS_type.F₀ : (f = zeros(FT, n, nSpec); f[1, :] .= 1; f)
    shift = Int32.(RS_type.i_λ₁λ₀); wR = Float64.(RS_type.ϖ_λ₁λ₀)
    d_τr, d_τa, d_F0, d_Zpp, d_Zmp = CuArray{FT}(τ_rayl), CuArray{FT}(τ_abs), CuArray{FT}(F0), CuArray(Zpp), CuArray(Zmp)
    d_ZppR, d_ZmpR = CuArray(ZppR), CuArray(ZmpR)
    R_d = CuArray{FT}(undef, nV, n, nSpec); T_d = similar(R_d); ieR_d = similar(R_d); ieT_d = similar(R_d)
    GC.@preserve μh wh ω ft τa vw ndoubl scatter iface rz vstream ρ shift wR d_τr d_τa d_F0 d_Zpp d_Zmp d_ZppR d_ZmpR begin
        d = ColumnDesc(sizeof(ColumnDesc), N, n, nSpec, Nz, nAer, nV, nM, iμ₀, lamb ? 0 : 1,
            μ₀, lamb ? brdf0.albedo : 0.0, RS_type.ϖ_Cabannes[iB],
            pointer(μh), pointer(wh), pointer(ω), pointer(ft), pointer(τa), pointer(vw),
            pointer(ndoubl), pointer(scatter), pointer(iface), pointer(rz), pointer(vstream),
            _p(d_τr), _p(d_τa), _p(d_F0), _p(d_Zpp), _p(d_Zmp), lamb ? CU_NULL : _p(ρ), Ptr{Float64}(C_NULL))
        # one device owns the whole band: owner range = 1:nSpec (a multi-GPU driver passes one contiguous range per
        # device; the library recomputes the elastic halo of the range locally, no communication)
        rd = RrsDesc(sizeof(RrsDesc), length(shift), 0, nSpec, pointer(shift), pointer(wR), _p(d_ZppR), _p(d_ZmpR))
        _check(ccall((:vsm_rt_column_rrs_f64, LIB), Cint,
            (Ptr{Cvoid}, Ref{ColumnDesc}, Ref{RrsDesc}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}),
            CTX[], d, rd, R_d, T_d, ieR_d, ieT_d))
        Architectures.synchronize_if_gpu()
    end
    hdr = zeros(FT, nV, n, nSpec)                              # interaction_hdrf! outputs are not produced on this path
    return Array(R_d), Array(T_d), Array(ieR_d), Array(ieT_d), hdr, zeros(FT, nSpec), zeros(FT, nSpec)   # rt_run.jl:532
end

# ---- single scattering: rt_run_ss(::noRS, model, iBand) (rt_run.jl:593-755) is ONE call on the same ColumnDesc:
#   ccall((:vsm_rt_column_ss_f64, LIB), Cint, (Ptr{Cvoid}, Ref{ColumnDesc}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}),
#         CTX[], d, R_d, T_d, hemR_d, hemT_d)        # R_SFI/T_SFI (nVZA,n,nSpec), hem_R/hem_T [nSpec]
# with `d` filled exactly as in rt_run(::noRS) above (ndoubl / scatter / iface are not read by this driver); the method
#   CoreRT.rt_run_ss(RS_type::CoreRT.noRS{FT}, model::CoreRT.RTModel{<:Architectures.GPU,FT}, iBand; kwargs...)
# returns (Array(R_d), Array(T_d), zeros, zeros, Array(hemR_d), Array(hemT_d)) -- rt_run.jl:754.
#
# ---- LinMode: vsm_rt_column_lin_f64 takes the same ColumnDesc plus
#   mutable struct LinDesc; struct_size::Int32; n_params::Int32; dtau_aer::Ptr{Float64}; dalbedo::Ptr{Float64};
#                           dtau_rayl::CuPtr{Cvoid}; dtau_abs::CuPtr{Cvoid}; end
# filled from lin_model.τ̇_abs / τ̇_aer (types.jl OpticsLin) for the parameters that act through optical depths and the
# albedo; outputs Ṙ, Ṫ (nVZA, nStokes, nSpec, Nparams) as rt_run_lin.jl:185-190.  vsmom_b200/rt_run.py::rt_run_lin is
# the executable form.
#
# ---- Multi-GPU from one Julia process: vsm_multi_create(ctxs, devices, ndev) + vsm_rt_column_sharded_f64(ctxs, ndev,
# desc with HOST spectral arrays, host outputs) shard the spectral axis over the devices inside the library (one host
# thread per device, no data-path collective, outputs land in the caller's host arrays); see INTEGRATION.md §4.
end # module

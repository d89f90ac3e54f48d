"""Host-side model for the B200 CoreRT path: the arrays that the reference's
``model_from_parameters`` hands to ``rt_run`` (src/CoreRT/tools/model_from_parameters.jl:211-524),
built here with NumPy.  This is setup code that runs once per scene; the
per-Fourier-moment / per-layer hot loop runs in the CUDA library.

Names mirror the reference: ``QuadPoints`` (src/CoreRT/types.jl:872-891),
``GreekCoefs``, ``rt_set_streams`` (src/CoreRT/tools/rt_set_streams.jl:25-47),
``compute_Z_moments`` (src/Scattering/compute_Z_matrices.jl:26-110),
``get_greek_rayleigh`` (src/Scattering/mie_helper_functions.jl:454-468).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

POL_N = {"Stokes_I": 1, "Stokes_IQ": 2, "Stokes_IQU": 3, "Stokes_IQUV": 4}


def _cosd(x: float) -> float:
    r = math.fmod(float(x), 360.0)
    r = r + 360.0 if r < 0 else r
    table = {0.0: 1.0, 60.0: 0.5, 90.0: 0.0, 120.0: -0.5, 180.0: -1.0, 240.0: -0.5, 270.0: 0.0, 300.0: 0.5}
    return table[r] if r in table else math.cos(math.radians(r))


def _sind(x: float) -> float:
    r = math.fmod(float(x), 360.0)
    r = r + 360.0 if r < 0 else r
    table = {0.0: 0.0, 30.0: 0.5, 90.0: 1.0, 150.0: 0.5, 180.0: 0.0, 210.0: -0.5, 270.0: -1.0, 330.0: -0.5}
    return table[r] if r in table else math.sin(math.radians(r))


@dataclass
class QuadPoints:
    """src/CoreRT/types.jl:872-891 (``iμ₀`` is 1-based as in Julia)."""
    mu0: float
    imu0: int
    imu0Nstart: int
    qp_mu: np.ndarray
    wt_mu: np.ndarray
    qp_muN: np.ndarray
    wt_muN: np.ndarray
    Nquad: int
    Nstreams: int


def rt_set_streams(nstreams: int, sza: float, vza: Sequence[float], pol_n: int, FT=np.float64) -> QuadPoints:
    """GaussLegQuad streams: ``nstreams`` weighted half-space nodes plus the
    zero-weight VZA / SZA output nodes (rt_set_streams.jl:25-47,153-161)."""
    x, w = np.polynomial.legendre.leggauss(nstreams)
    nodes = [FT(0.5 * v + 0.5) for v in x]
    weights = [FT(0.5 * v) for v in w]
    mu0 = _cosd(sza)
    seen, qp = set(), []
    for v in nodes + [FT(_cosd(z)) for z in vza] + [FT(mu0)]:
        if v not in seen:
            seen.add(v)
            qp.append(v)
    qp_mu = np.asarray(qp, dtype=FT)
    wt_mu = np.zeros(len(qp), FT)
    wt_mu[:nstreams] = weights
    imu0 = int(np.argmin(np.abs(qp_mu - FT(mu0)))) + 1
    return QuadPoints(float(FT(mu0)), imu0, pol_n * (imu0 - 1) + 1, qp_mu, wt_mu,
                      np.repeat(qp_mu, pol_n), np.repeat(wt_mu, pol_n), len(qp),
                      int(np.count_nonzero(wt_mu)))


@dataclass
class GreekCoefs:
    alpha: np.ndarray
    beta: np.ndarray
    gamma: np.ndarray
    delta: np.ndarray
    epsilon: np.ndarray
    zeta: np.ndarray


def get_greek_rayleigh(depol: float) -> GreekCoefs:
    p = (1 - depol) / (1 + depol / 2)
    r = (1 - 2 * depol) / (1 - depol)
    z3 = np.zeros(3)
    return GreekCoefs(np.array([0, 0, 3 * p]), np.array([1, 0, 0.5 * p]),
                      np.array([0, 0, p * math.sqrt(1.5)]), np.array([0, 1.5 * p * r, 0]), z3.copy(), z3.copy())


def _prt(mu: np.ndarray, lmax: int, m: int):
    """Normalised generalised spherical functions P_l^m, R_l^m, T_l^m for one
    Fourier order ``m`` and l = 0..lmax-1, vectorised over mu
    (recurrences of src/Scattering/legendre_functions.jl:24-185; returns the
    "true" T, i.e. minus the array that routine fills)."""
    n = mu.size
    P = np.zeros((lmax, n)); R = np.zeros((lmax, n)); Tn = np.zeros((lmax, n))
    c = mu
    s = np.sqrt(1.0 - c * c)
    sq = math.sqrt
    for l in range(m, lmax):
        if m == 0:
            if l == 0:
                P[l] = 1.0
            elif l == 1:
                P[l] = c
            elif l == 2:
                P[l] = 0.5 * (3.0 * c * c - 1.0)
                R[l] = 0.5 * sq(1.5) * s * s
            else:
                P[l] = (P[l - 1] * (2 * l - 1) * c - P[l - 2] * (l - 1)) / l
                R[l] = (R[l - 1] * (2 * l - 1) * c - R[l - 2] * sq((l + 1) * (l - 3))) / sq(l * l - 4)
        elif m == 1 and l == 1:
            P[l] = sq(0.5) * s
        elif m == 1 and l == 2:
            m1 = sq(1 / 6)
            P[l] = m1 * 3 * c * s
            R[l] = -m1 * c * sq(1.5) * s
            Tn[l] = m1 * sq(1.5) * s
        elif l == m:            # m >= 2 seed
            f1 = np.ones(n); f2 = np.ones(n)
            half = s / 2
            for k in range(1, m + 1):
                f1 = f1 * ((2 * k - 1) * s) / sq(k * (k + m))
                f2 = f2 * half * (sq((m + k) / (k - 2)) if k > 2 else 1.0)
            big = s > 1e-8
            s2 = np.where(big, s * s, 1.0)
            aii = np.where(big, f2 * (1.0 + c * c) / s2, 0.5 if m == 2 else 0.0)
            aij = np.where(big, f2 * (2 * c) / s2, 0.5 if m == 2 else 0.0)
            P[l] = f1
            R[l] = aii
            Tn[l] = -aij
        else:
            if m == 1:
                m1 = sq((l - 1) / (l + 1))
                m2 = m1 * sq((l - 2) / l)
            elif l == m + 1:
                m1 = sq(1 / (l + m))
                m2 = 0.0
            else:
                m1 = sq((l - m) / (l + m))
                m2 = m1 * sq((l - m - 1) / (l + m - 1))
            Pm2 = P[l - 2] if l - 2 >= 0 else 0.0
            Rm2 = R[l - 2] if l - 2 >= 0 else 0.0
            Tm2 = Tn[l - 2] if l - 2 >= 0 else 0.0
            P[l] = (m1 * P[l - 1] * (2 * l - 1) * c - m2 * Pm2 * (l - 1 + m)) / (l - m)
            Z = (2 * m * (2 * l - 1)) / (l * (l - 1))
            Y = ((l + m - 1) / (l - 1)) * sq((l - 3) * (l + 1)) if (l - 3) * (l + 1) >= 0 else 0.0
            X = ((l - m) / l) * sq(l * l - 4)
            R[l] = (m1 * R[l - 1] * (2 * l - 1) * c - m2 * Rm2 * Y + m1 * Tn[l - 1] * Z) / X
            Tn[l] = (m1 * Tn[l - 1] * (2 * l - 1) * c - m2 * Tm2 * Y + m1 * R[l - 1] * Z) / X
    return P, R, -Tn


def _prt_pm(mu: np.ndarray, lmax: int, m: int):
    """``(_prt(mu), _prt(-mu))`` in one pass over l (the recurrences are element-wise in mu)."""
    n = mu.size
    P, R, T = _prt(np.concatenate([mu, -mu]), lmax, m)
    return (P[:, :n], R[:, :n], T[:, :n]), (P[:, n:], R[:, n:], T[:, n:])


def compute_Z_moments(pol_n: int, mu: np.ndarray, greek: GreekCoefs, m: int, prt=None):
    """Fourier moment ``m`` (0-based) of the phase matrix on the stream cosines:
    returns ``(Z⁺⁺, Z⁻⁺)`` of shape (N, N), N = len(mu)*pol_n, float64
    (compute_Z_matrices.jl:26-110).  ``prt``: ``_prt_pm(mu, L, m)`` with L >= len(greek.beta),
    when the caller has them already (they depend on the streams and m only, not on the scatterer)."""
    mu = np.asarray(mu, dtype=np.float64)
    nmu = mu.size
    lmax = len(greek.beta)
    fact = 0.5 if m == 0 else 1.0
    if prt is None:
        (Pp, Rp, Tp), (Pn, Rn, Tn) = _prt_pm(mu, lmax, m)
    else:
        (Pp, Rp, Tp), (Pn, Rn, Tn) = ((x[:lmax] for x in side) for side in prt)

    def pi_mats(P, R, T):
        Pi = np.zeros((lmax, nmu, pol_n, pol_n))
        Pi[:, :, 0, 0] = P
        if pol_n >= 2:
            Pi[:, :, 1, 1] = R
        if pol_n >= 3:
            Pi[:, :, 1, 2] = -T
            Pi[:, :, 2, 1] = -T
            Pi[:, :, 2, 2] = R
        if pol_n == 4:
            Pi[:, :, 3, 3] = P
        return Pi

    B = np.zeros((lmax, 4, 4))
    B[:, 0, 0] = greek.beta; B[:, 0, 1] = greek.gamma; B[:, 1, 0] = greek.gamma
    B[:, 1, 1] = greek.alpha; B[:, 2, 2] = greek.zeta; B[:, 2, 3] = greek.epsilon
    B[:, 3, 2] = -np.asarray(greek.epsilon); B[:, 3, 3] = greek.delta
    B = B[:, :pol_n, :pol_n]
    Pi_p = pi_mats(Pp, Rp, Tp)[m:]
    Pi_n = pi_mats(Pn, Rn, Tn)[m:]
    N = nmu * pol_n
    nl = Pi_p.shape[0]
    # sum over (l, c) as one matrix product: (i a | l c) (l c | j d)
    left = np.matmul(Pi_p, B[m:, None]).transpose(1, 2, 0, 3).reshape(N, nl * pol_n)
    App = (left @ Pi_p.transpose(0, 2, 1, 3).reshape(nl * pol_n, N)).reshape(nmu, pol_n, nmu, pol_n)
    Amp = (left @ Pi_n.transpose(0, 2, 1, 3).reshape(nl * pol_n, N)).reshape(nmu, pol_n, nmu, pol_n)
    sgn = np.ones((pol_n, pol_n))
    for i in range(pol_n):
        for j in range(pol_n):
            if (i <= 1 and j >= 2) or (i >= 2 and j <= 1):
                sgn[i, j] = -1.0
    Zpp = (2 * fact * App).reshape(N, N)
    Zmp = (2 * fact * Amp * sgn[None, :, None, :]).reshape(N, N)
    return Zpp, Zmp


@dataclass
class AerosolOptics:
    """Scattering.AerosolOptics: Greek coefficients, ω̃, truncation factor fᵗ."""
    greek_coefs: GreekCoefs
    omega: float
    ft: float = 0.0


@dataclass
class RTNumericalParameters:
    """src/CoreRT/types.jl:713-756"""
    dtau_max_threshold: float = 1e-3
    dtau_min_floor: Optional[float] = None


@dataclass
class RTModel:
    """The slice of the reference's ``RTModel`` that the noRS hot path reads."""
    polarization_type: str
    quad_points: QuadPoints
    sza: float
    vza: Sequence[float]
    vaz: Sequence[float]
    tau_rayl: np.ndarray            # (nSpec, Nz)   model.τ_rayl[iB]
    tau_abs: np.ndarray             # (nSpec, Nz)   model.τ_abs[iB]
    tau_aer: np.ndarray             # (nAer, Nz)    model.τ_aer[iB]
    aerosol_optics: List[AerosolOptics]
    greek_rayleigh: GreekCoefs
    albedo: float                   # LambertianSurfaceScalar
    m_max: int
    varpi_cabannes: float = 1.0
    numerics: RTNumericalParameters = field(default_factory=RTNumericalParameters)
    float_type: type = np.float64
    spec_grid: Optional[np.ndarray] = None
    # BRDF surface (any AbstractSurfaceType of the reference other than the scalar Lambertian): Fourier
    # moments of R_surf, shape (m_max+1, N, N) as built by create_surface_layer! before the
    # Diagonal(qp_muN.*wt_muN) scaling (rpv_surface.jl:69-75).  None -> Lambertian(albedo).
    surface_rho: Optional[np.ndarray] = None
    # Cox-Munk TMS correction (apply_ss_correction!, rt_run.jl:520-524): (nVZA, pol_n) geometry factor from
    # coxmunk.coxmunk_ss_correction(); None for every other surface
    ss_correction: Optional[np.ndarray] = None
    # extra sources of the reference's source-term seam (rt_run(model; sources = ...)):
    #   thermal_B (nSpec, Nz): ThermalEmission.B_layer (Planck radiance per layer; Sources/thermal_emission.jl:92-115)
    #   sif0 (nSpec, pol_n):   SurfaceSIF.SIF0 (surface fluorescence irradiance; Sources/surface_sif.jl:153-178)
    thermal_B: Optional[np.ndarray] = None
    sif0: Optional[np.ndarray] = None

    @property
    def pol_n(self) -> int:
        return POL_N[self.polarization_type]

    @property
    def nSpec(self) -> int:
        return self.tau_abs.shape[0]

    @property
    def Nz(self) -> int:
        return self.tau_abs.shape[1]


def derive_m_max(aerosols: Sequence[AerosolOptics], nstreams: int) -> int:
    """Rayleigh (2), aerosol (len(β)-1), Lambertian surface (0), clamped to the
    stream cap 2·nstreams-1 (src/CoreRT/component_m_max.jl:51-117)."""
    m = 2
    for a in aerosols:
        m = max(m, len(a.greek_coefs.beta) - 1)
    return max(min(m, 2 * nstreams - 1), 0)


# ---------------------------------------------------------------------------
# control flow that the reference derives from FULL-band reductions
# ---------------------------------------------------------------------------
def doubling_number(dtau_max, tau_end, FT):
    """src/CoreRT/tools/rt_helper_functions.jl:49-69 -> ndoubl"""
    dtau_max = FT(dtau_max); tau_end = FT(tau_end)
    if tau_end <= dtau_max:
        return 0
    q1 = np.log10(FT(2)); q2 = np.log10(dtau_max); q3 = np.log10(tau_end)
    tlimit = FT((q3 - q2) / q1)
    nlimit = int(np.floor(tlimit))
    if FT(tlimit - FT(nlimit)) < np.finfo(FT).eps:
        return nlimit
    return nlimit + 1


@dataclass
class LayerControl:
    ndoubl: np.ndarray
    scatter: np.ndarray
    iface: np.ndarray
    rayl_allzero: np.ndarray


def layer_control(model: RTModel) -> LayerControl:
    """ndoubl / scatter / ScatteringInterface per layer from maximum(τ.*ϖ) over the
    whole band (rt_kernel.jl:197,266-287; compEffectiveLayerProperties.jl:75-93;
    rt_helper_functions.jl:15-33).  Must be evaluated on the full band before the
    spectral axis is sharded across GPUs."""
    FT = model.float_type
    eps = np.finfo(FT).eps
    q = model.quad_points
    thr = FT(model.numerics.dtau_max_threshold)
    flo = FT(1024 * eps if model.numerics.dtau_min_floor is None else model.numerics.dtau_min_floor)
    real = q.qp_mu[q.wt_mu > eps]
    mu_min = FT(real.min() if real.size else q.qp_mu.min())
    Nz = model.Nz
    nd = np.zeros(Nz, np.int32); sc = np.zeros(Nz, np.int32)
    itf = np.zeros(Nz, np.int32); rz = np.zeros(Nz, np.int32)
    prev = 0
    for iz in range(Nz):
        tau = model.tau_rayl[:, iz].astype(FT)
        varpi = np.full_like(tau, FT(model.varpi_cabannes))
        rz[iz] = int(np.all(tau * varpi == 0))
        for ia, a in enumerate(model.aerosol_optics):
            ft, om = FT(a.ft), FT(a.omega)
            ty = FT((FT(1) - ft * om) * FT(model.tau_aer[ia, iz]))
            py = FT((FT(1) - ft) * om / (FT(1) - ft * om))
            w = tau * varpi + ty * py
            tau = tau + ty
            varpi = w / np.where(tau > 0, tau, FT(1))
        wx = tau * varpi
        tau = tau + model.tau_abs[:, iz].astype(FT)
        varpi = wx / np.where(tau > 0, tau, FT(1))
        mx = FT(np.max(tau * varpi))
        s = bool(mx > 2 * eps)
        if iz == 0:
            prev = 3 if s else 0
        elif prev == 0:
            prev = 1 if s else 0
        else:
            prev = 3 if s else 2
        sc[iz] = int(s); itf[iz] = prev
        if s:
            dtau_max = max(flo, min(mx, FT(thr * mu_min)))
            nd[iz] = doubling_number(dtau_max, mx, FT)
    return LayerControl(nd, sc, itf, rz)


def vza_tables(model: RTModel):
    """``_precompute_vza_weights`` (postprocessing_vza.jl:23-37) for every moment:
    returns (vza_stream [nVZA] 1-based, weights (n_moments, nVZA, pol_n))."""
    FT = model.float_type
    q = model.quad_points
    n = model.pol_n
    nV = len(model.vza)
    stream = np.array([int(np.argmin(np.abs(q.qp_mu - FT(_cosd(v))))) + 1 for v in model.vza], np.int32)
    w = np.zeros((model.m_max + 1, nV, n))
    for m in range(model.m_max + 1):
        weight = float(FT(0.5 / math.pi)) if m == 0 else float(FT(1.0 / math.pi))
        for i in range(nV):
            c, s = _cosd(m * model.vaz[i]), _sind(m * model.vaz[i])
            w[m, i, :] = [weight * v for v in (c, c, s, s)[:n]]
    return stream, w


# ---------------------------------------------------------------------------
# synthetic O2-A-band scene: no HITRAN / Mie needed
# ---------------------------------------------------------------------------
DEFAULT_P_HALF = np.array([
    0.14, 0.22, 0.30, 0.39, 0.53, 0.71, 0.96, 1.28, 1.70, 2.27, 3.03, 4.03, 5.44, 7.26, 9.67,
    12.90, 17.23, 23.30, 31.00, 42.07, 56.09, 74.78, 99.69, 131.00, 176.85, 236.64, 314.58,
    418.87, 557.76, 735.00, 800.12, 849.00, 912.00, 980.00, 1005.0])   # DefaultParameters.yaml:87-92
DEFAULT_VZA = [60, 45, 30, 15, 0, 15, 30, 45, 60]                      # DefaultParameters.yaml:67-69
DEFAULT_VAZ = [180, 180, 180, 180, 0, 0, 0, 0, 0]


def bodhaine_tau_rayleigh(wl_um: np.ndarray, psurf_hpa: float) -> np.ndarray:
    """Bodhaine (1999) Eq. 30 column Rayleigh optical depth (atmo_prof.jl:251-256)."""
    l2 = wl_um ** 2
    t = 0.002152 * (1.0455996 - 341.29061 / l2 - 0.90230850 * l2) / (1 + 0.0027059889 / l2 - 85.968563 * l2)
    return t * (psurf_hpa / 1013.25)


def hg_greek(g: float, nterms: int) -> GreekCoefs:
    """Henyey-Greenstein-like Greek set: β_l=(2l+1)g^l; α, ζ, δ follow β for
    l>=2 (δ from l>=0); γ = ε = 0."""
    l = np.arange(nterms)
    beta = (2 * l + 1) * g ** l
    a = np.where(l >= 2, beta, 0.0)
    return GreekCoefs(a, beta, np.zeros(nterms), beta.copy(), np.zeros(nterms), a.copy())


def _synthetic_absorption(nu: np.ndarray, p_half: np.ndarray, seed: int, nu_lo: float, nu_hi: float,
                          col_max: float, nlines: int = 60) -> np.ndarray:
    """(nSpec, Nz) absorption optical depths of `nlines` pressure-broadened Lorentz pseudo-lines with centres
    uniform in [nu_lo, nu_hi] and strengths log-uniform over four decades, scaled to a largest column optical
    depth `col_max`, plus a 1e-4 continuum."""
    dp = np.diff(p_half)
    pbar = 0.5 * (p_half[1:] + p_half[:-1])
    rng = np.random.default_rng(seed)
    nu_l = rng.uniform(nu_lo, nu_hi, nlines)
    a_l = 10.0 ** rng.uniform(-3.0, 1.0, nlines)
    wz = dp * pbar
    wz = wz / wz.sum()
    tau_abs = np.zeros((nu.size, len(dp)))
    for iz in range(len(dp)):
        gam = 0.05 * pbar[iz] / 1013.0 + 0.002
        lor = (gam / math.pi) / ((nu[:, None] - nu_l[None, :]) ** 2 + gam * gam)
        tau_abs[:, iz] = wz[iz] * (lor * a_l[None, :]).sum(axis=1)
    col = tau_abs.sum(axis=1)
    tau_abs *= col_max / max(col.max(), 1e-300)
    tau_abs += 1e-4 * wz[None, :]
    return tau_abs


def synthetic_o2a_model(nSpec: int = 10000, nstreams: int = 10, polarization_type: str = "Stokes_IQU",
                        aerosol: bool = True, float_type=np.float64, sza: float = 60.0,
                        vza=DEFAULT_VZA, vaz=DEFAULT_VAZ, albedo: float = 0.15,
                        p_half: np.ndarray = DEFAULT_P_HALF, depol: float = 0.0279,
                        seed: int = 20251017) -> RTModel:
    FT = float_type
    pol_n = POL_N[polarization_type]
    q = rt_set_streams(nstreams, sza, vza, pol_n, FT)
    nu = 12870.01 + 0.015 * np.arange(nSpec)
    wl = 1e4 / nu
    dp = np.diff(p_half)
    pbar = 0.5 * (p_half[1:] + p_half[:-1])
    tau_R = bodhaine_tau_rayleigh(wl, p_half[-1])
    tau_rayl = tau_R[:, None] * (dp / dp.sum())[None, :]
    tau_abs = _synthetic_absorption(nu, p_half, seed, 12900.0, 13180.0, 100.0)
    aerosols, tau_aer = [], np.zeros((0, len(dp)))
    if aerosol:
        aerosols = [AerosolOptics(hg_greek(0.7, 2 * nstreams), 0.95, 0.0)]
        prof = np.exp(-0.5 * ((pbar - 900.0) / 50.0) ** 2) * dp
        tau_aer = (0.3 * prof / prof.sum())[None, :]
    return RTModel(polarization_type, q, sza, list(vza), list(vaz), tau_rayl.astype(FT), tau_abs.astype(FT),
                   tau_aer, aerosols, get_greek_rayleigh(depol), albedo, derive_m_max(aerosols, nstreams),
                   1.0, RTNumericalParameters(), FT, nu)


# config/ocean_coxmunk.yaml: atmospheric_profile.p (34 half levels -> 33 layers), geometry
OCEAN_P_HALF = np.array([0.14, 0.22, 0.30, 0.39, 0.53, 0.71, 0.96, 1.28, 1.70, 2.27, 3.03, 4.03, 5.44, 7.26, 9.67, 12.90,
                         17.23, 23.30, 31.00, 42.07, 56.09, 74.78, 99.69, 131.00, 176.85, 236.64, 314.58, 418.87,
                         557.76, 735.00, 800.12, 849.00, 912.00, 980.00])
OCEAN_VZA = (60.0, 45.0, 30.0, 15.0, 0.0, 15.0, 30.0, 45.0, 60.0)
OCEAN_VAZ = (180.0, 180.0, 180.0, 180.0, 0.0, 0.0, 0.0, 0.0, 0.0)


def synthetic_ocean_model(nSpec: int = 1000, nstreams: int = 20, polarization_type: str = "Stokes_IQU",
                          aerosol: bool = True, wind_speed: float = 5.0, float_type=np.float64,
                          ss_correction: bool = True, seed: int = 20251017) -> RTModel:
    """BASELINE.json configs[2]: the scene of ``config/ocean_coxmunk.yaml`` (geometry sza 30, the nine viewing
    angles, the 33-layer profile, a CoxMunkSurface at 5 m/s with whitecaps and shadowing, band around
    19417.5 cm^-1 = 515 nm) with the polarized + aerosol, 20-stream settings of the benchmark row: 20 + 5
    zero-weight nodes -> Nquad = 25 -> N = 75 (IQU) / 100 (IQUV), Fourier loop to 2*nstreams - 1 (Cox-Munk runs
    to the stream cap, component_m_max.jl:72).  The spectral axis is a synthetic nSpec-point grid across the band
    with pseudo-line absorption (column 1e-4 .. 3) so that layers differ spectrally."""
    from .coxmunk import CoxMunkSurface, coxmunk_ss_correction, coxmunk_surface_rho
    FT = float_type
    pol_n = POL_N[polarization_type]
    q = rt_set_streams(nstreams, 30.0, OCEAN_VZA, pol_n, FT)
    nu = 19417.0 + np.arange(nSpec) / max(nSpec, 1)
    wl = 1e4 / nu
    dp = np.diff(OCEAN_P_HALF)
    pbar = 0.5 * (OCEAN_P_HALF[1:] + OCEAN_P_HALF[:-1])
    tau_R = bodhaine_tau_rayleigh(wl, OCEAN_P_HALF[-1])
    tau_rayl = tau_R[:, None] * (dp / dp.sum())[None, :]
    tau_abs = _synthetic_absorption(nu, OCEAN_P_HALF, seed, 19416.8, 19418.2, 3.0, nlines=24)
    aerosols, tau_aer = [], np.zeros((0, len(dp)))
    if aerosol:
        aerosols = [AerosolOptics(hg_greek(0.7, 2 * nstreams), 0.95, 0.0)]
        prof = np.exp(-0.5 * ((pbar - 900.0) / 50.0) ** 2) * dp
        tau_aer = (0.3 * prof / prof.sum())[None, :]
    m_max = 2 * nstreams - 1
    surf = CoxMunkSurface(wind_speed)
    model = RTModel(polarization_type, q, 30.0, list(OCEAN_VZA), list(OCEAN_VAZ), tau_rayl.astype(FT), tau_abs.astype(FT),
                    tau_aer, aerosols, get_greek_rayleigh(0.0279), 0.0, m_max, 1.0, RTNumericalParameters(), FT, nu)
    model.surface_rho = coxmunk_surface_rho(q, pol_n, m_max, surf)
    if ss_correction:
        model.ss_correction = coxmunk_ss_correction(surf, pol_n, OCEAN_VZA, OCEAN_VAZ, float(q.mu0), m_max)
    return model


# ----------------------------------------------------------------------------
# BRDF surfaces: host-side Fourier moments (the reference computes them on the host too,
# rpv_surface.jl:166-196); the device consumes them through RTModel.surface_rho.
# ----------------------------------------------------------------------------
def brdf_fourier_moments(quad: QuadPoints, pol_n: int, m_max: int, rho, nquad: int = 100) -> np.ndarray:
    """Fourier moments 0..m_max of a scalar BRDF ``rho(mu_i, mu_r, dphi)`` (NumPy-broadcasting callable) as the
    reference's ``reflectance(brdf, pol_type, mu, m)`` builds them (rpv_surface.jl:166-196): 100-point
    Gauss-Legendre rule in azimuth over [0, pi], only the Stokes-I block is filled, times the overall factor 2
    (``ff`` for m > 0, the explicit ``2*`` of create_surface_layer! for m = 0, :69-73).  Shape (m_max+1, N, N),
    ready for ``RTModel.surface_rho``."""
    mu = np.asarray(quad.qp_mu, np.float64)
    x, w = np.polynomial.legendre.leggauss(nquad)
    phi = 0.5 * math.pi * (x + 1.0)
    w = 0.5 * math.pi * w
    r = rho(mu[:, None, None], mu[None, :, None], phi[None, None, :])          # (Nquad, Nquad, nquad)
    nq = mu.size
    out = np.zeros((m_max + 1, nq * pol_n, nq * pol_n))
    for m in range(m_max + 1):
        c = (r * (w * np.cos(m * phi))[None, None, :]).sum(axis=2)
        out[m, ::pol_n, ::pol_n] = 2.0 * c / math.pi
    return out


def rpv_reflectance(rho0: float, rho_c: float, k: float, theta_hs: float):
    """RPV (Rahman-Pinty-Verstraete) BRDF with the reference's sign conventions (rpv_surface.jl:113-144)."""
    def rho(mi, mr, dphi):
        ti, tr = np.arccos(mi), np.arccos(mr)
        cosg = -mi * mr + np.sin(ti) * np.sin(tr) * np.cos(dphi)
        G = np.sqrt(np.tan(ti) ** 2 + np.tan(tr) ** 2 + 2 * np.tan(ti) * np.tan(tr) * np.cos(dphi))
        M = (mi * mr) ** (k - 1) / (mi + mr) ** (1 - k)
        th = -theta_hs
        F = (1 - th * th) / (1 + th * th + 2 * th * cosg) ** 1.5
        H = 1 + (1 - rho_c) / (1 + G)
        return rho0 * M * F * H
    return rho


def rossli_reflectance(fiso: float, fvol: float, fgeo: float):
    """Ross-Li BRDF: isotropic + Ross-thick volumetric + Li-sparse geometric kernels with the RAMI defaults
    h/b = 2, b/r = 1 (rossli_surface.jl: reflectance, RossLi_K_vol, RossLi_K_geo)."""
    def rho(mi, mr, dphi):
        dphi = math.pi - dphi
        ti, tr = np.arccos(mi), np.arccos(mr)
        xi = np.arccos(np.clip(np.cos(ti) * np.cos(tr) + np.sin(ti) * np.sin(tr) * np.cos(dphi), -1.0, 1.0))
        kvol = ((math.pi / 2 - xi) * np.cos(xi) + np.sin(xi)) / (np.cos(ti) + np.cos(tr)) - math.pi / 4
        tip, trp = np.arctan(np.tan(ti) * 1.0), np.arctan(np.tan(tr) * 1.0)
        xip = np.arccos(np.clip(np.cos(tip) * np.cos(trp) + np.sin(tip) * np.sin(trp) * np.cos(dphi), -1.0, 1.0))
        D = np.sqrt(np.maximum(np.tan(tip) ** 2 + np.tan(trp) ** 2 - 2 * np.tan(tip) * np.tan(trp) * np.cos(dphi), 0.0))
        sec = 1 / np.cos(tip) + 1 / np.cos(trp)
        ct = np.clip(2.0 * np.sqrt(D ** 2 + (np.tan(tip) * np.tan(trp) * np.sin(dphi)) ** 2) / sec, -1.0, 1.0)
        t = np.arccos(ct)
        O = (1 / math.pi) * (t - np.sin(t) * np.cos(t)) * sec
        kgeo = O - sec + 0.5 * (1 + np.cos(xip)) / (np.cos(tip) * np.cos(trp))
        return fiso + fvol * kvol + fgeo * kgeo
    return rho


def rossli_surface_rho(quad: QuadPoints, pol_n: int, m_max: int, fiso: float, fvol: float, fgeo: float,
                       nquad: int = 100) -> np.ndarray:
    """Fourier moments of the Ross-Li reflectance matrix, shape (m_max+1, N, N)."""
    return brdf_fourier_moments(quad, pol_n, m_max, rossli_reflectance(fiso, fvol, fgeo), nquad)


def rpv_surface_rho(quad: QuadPoints, pol_n: int, m_max: int, rho0: float, rho_c: float, k: float,
                    theta_hs: float, nquad: int = 100) -> np.ndarray:
    """Fourier moments of the RPV reflectance matrix, shape (m_max+1, N, N)."""
    return brdf_fourier_moments(quad, pol_n, m_max, rpv_reflectance(rho0, rho_c, k, theta_hs), nquad)

"""
CPU oracle for the vSmartMOM CoreRT matrix-operator hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module.  The product path (``vsmartmom.jl_b200/``) never imports it and has
no CPU fallback.

This is a NumPy restatement (batch-first arrays ``(nSpec, N, N)`` instead of
Julia's column-major ``(N, N, nSpec)``; same arithmetic) of the reference's
pure-Julia algorithm.  The reference cannot be executed in this container (no
Julia toolchain), so the oracle is pinned against the reference's own golden
tables instead (tests/test_oracle_golden.py):

* Natraj (2009) polarized Rayleigh table   test/benchmarks/natraj_trues.jl
  asserted at                              test/test_CoreRT.jl:110-157
* 6SV1 scalar Rayleigh reflectances        test/benchmarks/6SV1_R_trues.jl
  asserted at                              test/test_CoreRT.jl:7-43
* Siewert (2000) problem IIA / VLIDORT     test/vlidort_baseline/reference_data/
  asserted at                              test/vlidort_baseline/cases/case_A_siewert2000.jl

Each function cites the reference file:line it follows (paths relative to
/root/reference).  Per-slice dense algebra is ``np.matmul`` (BLAS gemm per
slice) and ``np.linalg.inv`` (LAPACK getrf+getri per slice) -- the same
family as the reference CPU path src/CoreRT/tools/cpu_batched.jl:32-37,65-82.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def cosd(x: float) -> float:
    """Julia's cosd: exact at multiples of 30/45/60/90 degrees where the
    result is rational."""
    r = math.fmod(float(x), 360.0)
    if r < 0:
        r += 360.0
    exact = {0.0: 1.0, 60.0: 0.5, 90.0: 0.0, 120.0: -0.5, 180.0: -1.0,
             240.0: -0.5, 270.0: 0.0, 300.0: 0.5}
    if r in exact:
        return exact[r]
    return math.cos(math.radians(r))


def sind(x: float) -> float:
    r = math.fmod(float(x), 360.0)
    if r < 0:
        r += 360.0
    exact = {0.0: 0.0, 30.0: 0.5, 90.0: 1.0, 150.0: 0.5, 180.0: 0.0,
             210.0: -0.5, 270.0: -1.0, 330.0: -0.5}
    if r in exact:
        return exact[r]
    return math.sin(math.radians(r))


def _unique_preserve(x: Sequence[float]) -> np.ndarray:
    """Julia ``unique``: first occurrence order, exact equality."""
    seen = set()
    out = []
    for v in x:
        if v not in seen:
            seen.add(v)
            out.append(v)
    return np.asarray(out)


def nearest_point(f_array: np.ndarray, f: float) -> int:
    """src/CoreRT/tools/rt_helper_functions.jl:76 (0-based result)."""
    return int(np.argmin(np.abs(np.asarray(f_array) - f)))


# --------------------------------------------------------------------------
# quadrature   src/CoreRT/tools/rt_set_streams.jl:25-47
# --------------------------------------------------------------------------
@dataclass
class QuadPoints:
    mu0: float
    imu0: int            # 0-based index of the solar stream in qp_mu
    qp_mu: np.ndarray    # (Nquad,)
    wt_mu: np.ndarray    # (Nquad,)
    qp_muN: np.ndarray   # (Nquad*pol_n,)
    wt_muN: np.ndarray
    Nquad: int
    Nstreams: int
    pol_n: int


def gauleg(n: int, xmin: float, xmax: float):
    """src/Scattering/mie_helper_functions.jl:330-335 (FastGaussQuadrature
    gausslegendre mapped to [xmin, xmax])."""
    xi, w = np.polynomial.legendre.leggauss(n)
    xi = (xmax - xmin) / 2 * xi + (xmin + xmax) / 2
    w = w * (xmax - xmin) / 2
    return xi, w


def rt_set_streams_gauss(nstreams: int, sza: float, vza: Sequence[float],
                         pol_n: int, FT=np.float64) -> QuadPoints:
    """rt_set_streams(::GaussLegQuad, ...) src/CoreRT/tools/rt_set_streams.jl:25-47,153-161
    (kwarg form: Ltrunc = 2*nstreams-2 so (Ltrunc+2)//2 == nstreams)."""
    qp, wt = gauleg(nstreams, 0.0, 1.0)
    mu0 = cosd(sza)
    allmu = [FT(v) for v in qp] + [FT(cosd(v)) for v in vza] + [FT(mu0)]
    qp_mu = _unique_preserve(allmu).astype(FT)
    wt_mu = np.concatenate([wt.astype(FT), np.zeros(len(qp_mu) - len(wt), FT)])
    Nquad = len(qp_mu)
    Nstreams = int(np.count_nonzero(wt_mu))
    imu0 = nearest_point(qp_mu, FT(mu0))
    qp_muN = np.repeat(qp_mu, pol_n)
    wt_muN = np.repeat(wt_mu, pol_n)
    return QuadPoints(float(FT(mu0)), imu0, qp_mu, wt_mu, qp_muN, wt_muN, Nquad,
                      Nstreams, pol_n)


# --------------------------------------------------------------------------
# Greek coefficients and phase-matrix Fourier moments
# --------------------------------------------------------------------------
@dataclass
class GreekCoefs:
    alpha: np.ndarray
    beta: np.ndarray
    gamma: np.ndarray
    delta: np.ndarray
    epsilon: np.ndarray
    zeta: np.ndarray


def get_greek_rayleigh(depol: float) -> GreekCoefs:
    """src/Scattering/mie_helper_functions.jl:454-468"""
    dpl_p = (1 - depol) / (1 + depol / 2)
    dpl_r = (1 - 2 * depol) / (1 - depol)
    a = np.array([0.0, 0.0, 3 * dpl_p])
    b = np.array([1.0, 0.0, 0.5 * dpl_p])
    g = np.array([0.0, 0.0, dpl_p * math.sqrt(1.5)])
    d = np.array([0.0, dpl_p * dpl_r * 1.5, 0.0])
    e = np.zeros(3)
    z = np.zeros(3)
    return GreekCoefs(a, b, g, d, e, z)


def compute_associated_legendre_PRT(mu: np.ndarray, Lmax: int):
    """src/Scattering/legendre_functions.jl:24-185.  Returns (P, R, -T) with
    shape (len(mu), Lmax, Lmax), index [imu, l, m] (0-based l, m)."""
    mu = np.asarray(mu, dtype=np.float64)
    nmu = len(mu)
    P = np.zeros((nmu, Lmax, Lmax))
    R = np.zeros((nmu, Lmax, Lmax))
    T = np.zeros((nmu, Lmax, Lmax))
    sq = math.sqrt
    for m in range(Lmax):
        for l in range(m, Lmax):
            for i in range(nmu):
                cmu = mu[i]
                smu = sq(1.0 - cmu * cmu)
                if m == 0:
                    if l == 0:
                        P[i, l, m] = 1
                    elif l == 1:
                        P[i, l, m] = cmu
                    elif l == 2:
                        P[i, l, m] = 0.5 * (3.0 * cmu * cmu - 1.0)
                        R[i, l, m] = 0.5 * sq(1.5) * smu * smu
                    else:
                        P[i, l, m] = (P[i, l - 1, m] * (2 * l - 1) * cmu
                                      - P[i, l - 2, m] * (l - 1)) / l
                        Y = sq((l + 1) * (l - 3))
                        X = sq(l * l - 4)
                        R[i, l, m] = (R[i, l - 1, m] * (2 * l - 1) * cmu
                                      - R[i, l - 2, m] * Y) / X
                elif m == 1:
                    if l == 1:
                        P[i, l, m] = sq(0.5) * smu
                    elif l == 2:
                        m1 = sq(1 / 6)
                        cA = 3 * cmu * smu
                        cB = sq(1.5) * smu
                        P[i, l, m] = m1 * cA
                        R[i, l, m] = -m1 * cmu * cB
                        T[i, l, m] = m1 * cB
                    else:
                        m1 = sq((l - 1) / (l + 1))
                        m2 = m1 * sq((l - 2) / l)
                        Y = (l - 1 + m)
                        X = (l - m)
                        P[i, l, m] = (m1 * P[i, l - 1, m] * (2 * l - 1) * cmu
                                      - m2 * P[i, l - 2, m] * Y) / X
                        Z = (2 * m * (2 * l - 1)) / (l * (l - 1))
                        Y = ((l + m - 1) / (l - 1)) * sq((l - 3) * (l + 1))
                        X = ((l - m) / l) * sq(l * l - 4)
                        R[i, l, m] = (m1 * R[i, l - 1, m] * (2 * l - 1) * cmu
                                      - m2 * R[i, l - 2, m] * Y
                                      + m1 * T[i, l - 1, m] * Z) / X
                        T[i, l, m] = (m1 * T[i, l - 1, m] * (2 * l - 1) * cmu
                                      - m2 * T[i, l - 2, m] * Y
                                      + m1 * R[i, l - 1, m] * Z) / X
                else:
                    if l == m:
                        fact1 = 1.0
                        fact2 = 1.0
                        sfull = smu
                        shalf = sfull / 2
                        for k in range(1, m + 1):
                            fact1 = fact1 * ((2 * k - 1) * sfull) / sq(k * (k + m))
                            if k > 2:
                                fact2 = fact2 * shalf * sq((m + k) / (k - 2))
                            else:
                                fact2 = fact2 * shalf
                        if smu > 1e-8:
                            Km = fact2
                            Aii = Km * (1.0 + cmu * cmu) / (smu * smu)
                            Aij = Km * (2 * cmu) / (smu * smu)
                        else:
                            if m == 2:
                                Aii = 0.5
                                Aij = 0.5
                            else:
                                Aii = 0.0
                                Aij = 0.0
                        P[i, l, m] = fact1
                        R[i, l, m] = Aii
                        T[i, l, m] = -Aij
                    elif l == m + 1:
                        m1 = sq(1 / (l + m))
                        X = (l - m)
                        P[i, l, m] = (m1 * P[i, l - 1, m] * (2 * l - 1) * cmu) / X
                        Z = (2 * m * (2 * l - 1)) / (l * (l - 1))
                        X = ((l - m) / l) * sq(l * l - 4)
                        R[i, l, m] = (m1 * R[i, l - 1, m] * (2 * l - 1) * cmu
                                      + m1 * T[i, l - 1, m] * Z) / X
                        T[i, l, m] = (m1 * T[i, l - 1, m] * (2 * l - 1) * cmu
                                      + m1 * R[i, l - 1, m] * Z) / X
                    else:
                        m1 = sq((l - m) / (l + m))
                        m2 = m1 * sq((l - m - 1) / (l + m - 1))
                        Y = (l - 1 + m)
                        X = (l - m)
                        P[i, l, m] = (m1 * P[i, l - 1, m] * (2 * l - 1) * cmu
                                      - m2 * P[i, l - 2, m] * Y) / X
                        Z = (2 * m * (2 * l - 1)) / (l * (l - 1))
                        Y = ((l + m - 1) / (l - 1)) * sq((l - 3) * (l + 1))
                        X = ((l - m) / l) * sq(l * l - 4)
                        R[i, l, m] = (m1 * R[i, l - 1, m] * (2 * l - 1) * cmu
                                      - m2 * R[i, l - 2, m] * Y
                                      + m1 * T[i, l - 1, m] * Z) / X
                        T[i, l, m] = (m1 * T[i, l - 1, m] * (2 * l - 1) * cmu
                                      - m2 * T[i, l - 2, m] * Y
                                      + m1 * R[i, l - 1, m] * Z) / X
    return P, R, -T


def _construct_B(pol_n: int, g: GreekCoefs, l: int) -> np.ndarray:
    """src/Scattering/mie_helper_functions.jl:593-614 (l 0-based here)."""
    a, b, c, d, e, z = (g.alpha[l], g.beta[l], g.gamma[l], g.delta[l],
                        g.epsilon[l], g.zeta[l])
    if pol_n == 4:
        return np.array([[b, c, 0, 0], [c, a, 0, 0], [0, 0, z, e], [0, 0, -e, d]])
    if pol_n == 3:
        return np.array([[b, c, 0], [c, a, 0], [0, 0, z]])
    if pol_n == 2:
        return np.array([[b, c], [c, a]])
    return np.array([[b]])


def _construct_Pi(pol_n: int, P, R, T, l: int, m: int) -> np.ndarray:
    """src/Scattering/mie_helper_functions.jl:532-585 (sign_change=false).
    Returns (nmu, pol_n, pol_n)."""
    p, r, t = P[:, l, m], R[:, l, m], T[:, l, m]
    nmu = len(p)
    Pi = np.zeros((nmu, pol_n, pol_n))
    Pi[:, 0, 0] = p
    if pol_n >= 2:
        Pi[:, 1, 1] = r
    if pol_n >= 3:
        Pi[:, 1, 2] = -t
        Pi[:, 2, 1] = -t
        Pi[:, 2, 2] = r
    if pol_n == 4:
        Pi[:, 3, 3] = p
    return Pi


def compute_Z_moments(pol_n: int, mu: np.ndarray, greek: GreekCoefs, m: int,
                      FT=np.float64):
    """src/Scattering/compute_Z_matrices.jl:26-110.  ``m`` is the 0-based
    Fourier moment.  Returns (Zpp, Zmp), each (N, N) with N = len(mu)*pol_n."""
    mu = np.asarray(mu, dtype=np.float64)
    n = len(mu)
    fact = 0.5 if m == 0 else 1.0
    l_max = len(greek.beta)
    assert np.all((mu > 0) & (mu <= 1))
    P, R, T = compute_associated_legendre_PRT(mu, l_max)
    Pm, Rm, Tm = compute_associated_legendre_PRT(-mu, l_max)
    App = np.zeros((n, n, pol_n, pol_n))
    Amp = np.zeros((n, n, pol_n, pol_n))
    for l in range(m, l_max):
        B = _construct_B(pol_n, greek, l)
        Pi = _construct_Pi(pol_n, P, R, T, l, m)
        Pim = _construct_Pi(pol_n, Pm, Rm, Tm, l, m)
        left = Pi @ B                                   # (n, p, p)
        App += np.einsum('iab,jbc->ijac', left, Pi)
        Amp += np.einsum('iab,jbc->ijac', left, Pim)
    Zpp = np.zeros((n * pol_n, n * pol_n))
    Zmp = np.zeros((n * pol_n, n * pol_n))
    for i in range(pol_n):
        for j in range(pol_n):
            Zpp[i::pol_n, j::pol_n] = 2 * fact * App[:, :, i, j]
            if (i <= 1 and j >= 2) or (i >= 2 and j <= 1):
                Zmp[i::pol_n, j::pol_n] = -2 * fact * Amp[:, :, i, j]
            else:
                Zmp[i::pol_n, j::pol_n] = 2 * fact * Amp[:, :, i, j]
    return Zpp.astype(FT), Zmp.astype(FT)


# --------------------------------------------------------------------------
# Model container (the arrays model_from_parameters would produce)
# --------------------------------------------------------------------------
@dataclass
class Aerosol:
    greek: GreekCoefs
    omega: float          # ω̃
    ft: float = 0.0       # fᵗ (δ-truncation fraction)


@dataclass
class Model:
    pol_n: int
    quad: QuadPoints
    vza: Sequence[float]
    vaz: Sequence[float]
    tau_rayl: np.ndarray              # (nSpec, Nz)
    tau_abs: np.ndarray               # (nSpec, Nz)
    tau_aer: np.ndarray               # (nAer, Nz)
    aerosols: List[Aerosol]
    greek_rayleigh: GreekCoefs
    albedo: float
    m_max: int
    varpi_cabannes: float = 1.0
    dtau_max_threshold: float = 1e-3
    dtau_min_floor: Optional[float] = None   # None -> 1024*eps(FT)
    F0: Optional[np.ndarray] = None          # (pol_n, nSpec); default unit Stokes-I
    FT: type = np.float64
    # BRDF surface: Fourier moments of R_surf, (m_max+1, N, N) -- see brdf_surface(); None = Lambertian
    surface_rho: Optional[np.ndarray] = None
    # Cox-Munk TMS correction (apply_ss_correction!, rt_run.jl:520-524): (nVZA, pol_n) geometry factor from
    # coxmunk_ss_correction(); None = no correction (any surface other than CoxMunkSurface)
    ss_correction: Optional[np.ndarray] = None
    # extra sources (rt_run(model; sources = SourceSet(...))): ThermalEmission.B_layer as (nSpec, Nz) and
    # SurfaceSIF.SIF0 as (nSpec, pol_n); None = absent
    thermal_B: Optional[np.ndarray] = None
    sif0: Optional[np.ndarray] = None

    @property
    def nSpec(self):
        return self.tau_abs.shape[0]

    @property
    def Nz(self):
        return self.tau_abs.shape[1]


def derive_m_max(model_aerosols: List[Aerosol], nstreams: int) -> int:
    """src/CoreRT/component_m_max.jl:51-117 for Rayleigh + aerosols +
    Lambertian surface: max(2, len(beta)-1) clamped to 2*nstreams-1."""
    m = 2
    for a in model_aerosols:
        m = max(m, len(a.greek.beta) - 1)
    return max(min(m, 2 * nstreams - 1), 0)


# --------------------------------------------------------------------------
# doubling number   src/CoreRT/tools/rt_helper_functions.jl:49-69
# --------------------------------------------------------------------------
def doubling_number(dtau_max, tau_end, FT=np.float64):
    dtau_max = FT(dtau_max)
    tau_end = FT(tau_end)
    if tau_end <= dtau_max:
        return tau_end, 0
    q1 = np.log10(FT(2))
    q2 = np.log10(dtau_max)
    q3 = np.log10(tau_end)
    tlimit = FT((q3 - q2) / q1)
    nlimit = int(np.floor(tlimit))
    diff = FT(tlimit - FT(nlimit))
    if diff < np.finfo(FT).eps:
        return dtau_max, nlimit
    ndoubl = nlimit + 1
    x = FT(q3 - q1 * FT(ndoubl))
    return FT(10) ** x, ndoubl


def get_ndoubl(max_tau_varpi, quad: QuadPoints, FT, threshold=None, floor=None) -> int:
    """src/CoreRT/CoreKernel/rt_kernel.jl:266-287 (returns ndoubl only; the
    caller forms dτ = τ / 2^ndoubl)."""
    eps = np.finfo(FT).eps
    thr = FT(1e-3 if threshold is None else threshold)
    flo = FT(1024 * eps if floor is None else floor)
    real = quad.qp_mu[quad.wt_mu > eps]
    mu_min = FT(real.min() if real.size else quad.qp_mu.min())
    mx = FT(max_tau_varpi)
    dtau_max = max(flo, min(mx, FT(thr * mu_min)))
    _, nd = doubling_number(dtau_max, mx, FT)
    return nd


# --------------------------------------------------------------------------
# layer optical properties
#   src/CoreRT/LayerOpticalProperties/compEffectiveLayerProperties.jl:11-117
#   src/CoreRT/types.jl:1262-1308
# --------------------------------------------------------------------------
@dataclass
class CoreProps:
    tau: np.ndarray         # (nSpec,) or scalar
    varpi: np.ndarray       # (nSpec,) or scalar
    Zpp: np.ndarray         # (N, N) or (nSpec, N, N)
    Zmp: np.ndarray


def _add_scatt(x: CoreProps, y: CoreProps, with_Z: bool = True) -> CoreProps:
    """Base.:+ of two CoreScatteringOpticalProperties, types.jl:1262-1293.
    ``with_Z=False`` skips the Z mixing (only τ, ϖ are needed for control flow)."""
    tau = x.tau + y.tau
    wx = x.tau * x.varpi
    wy = y.tau * y.varpi
    w = wx + wy
    varpi = w / np.where(tau > 0, tau, np.ones_like(tau))
    if not with_Z:
        return CoreProps(tau, varpi, x.Zpp, x.Zmp)
    if np.all(wx == 0.0):
        return CoreProps(tau, varpi, y.Zpp, y.Zmp)
    if np.all(wy == 0.0):
        return CoreProps(tau, varpi, x.Zpp, x.Zmp)
    wy = np.atleast_1d(wy / w)
    wx = np.atleast_1d(wx / w)
    n = max(wx.size, wy.size)
    wx = np.broadcast_to(wx, (n,)).reshape(n, 1, 1)
    wy = np.broadcast_to(wy, (n,)).reshape(n, 1, 1)
    Zpp = wx * x.Zpp + wy * y.Zpp
    Zmp = wx * x.Zmp + wy * y.Zmp
    return CoreProps(tau, varpi, Zpp, Zmp)


def _add_abs(x: CoreProps, tau_abs: np.ndarray) -> CoreProps:
    """types.jl:1302-1308"""
    tau = x.tau + tau_abs
    wx = x.tau * x.varpi
    varpi = wx / np.where(tau > 0, tau, np.ones_like(tau))
    return CoreProps(tau, varpi, x.Zpp, x.Zmp)


def z_tables(model: Model, m: int):
    """Z-moment matrices of Rayleigh and every aerosol for Fourier moment m."""
    mu = model.quad.qp_mu
    return (compute_Z_moments(model.pol_n, mu, model.greek_rayleigh, m, model.FT),
            [compute_Z_moments(model.pol_n, mu, a.greek, m, model.FT) for a in model.aerosols])


def construct_core_optical_properties(model: Model, m: int, sel=slice(None), ztab=None, with_Z: bool = True):
    """constructCoreOpticalProperties, single band.  ``sel`` restricts the
    spectral axis (control-flow scalars must come from the full band; see
    band_control)."""
    FT = model.FT
    (Zpp_r, Zmp_r), aerZ = ztab if ztab is not None else z_tables(model, m)
    out = []
    for iz in range(model.Nz):
        combo = CoreProps(model.tau_rayl[sel, iz].astype(FT),
                          FT(model.varpi_cabannes), Zpp_r, Zmp_r)
        for ia, a in enumerate(model.aerosols):
            # createAero, compEffectiveLayerProperties.jl:67-72
            tau_mod = FT((1 - a.ft * a.omega) * model.tau_aer[ia, iz])
            varpi_mod = FT((1 - a.ft) * a.omega / (1 - a.ft * a.omega))
            combo = _add_scatt(combo, CoreProps(tau_mod, varpi_mod, *aerZ[ia]), with_Z)
        combo = _add_abs(combo, model.tau_abs[sel, iz].astype(FT))
        out.append(combo)
    return out


@dataclass
class BandControl:
    """Per-layer control flow derived from FULL-band reductions
    (rt_kernel.jl:197,282-283; compEffectiveLayerProperties.jl:75-93)."""
    scatter: List[bool]
    ndoubl: List[int]
    interface: List[int]     # 0=00, 1=01, 2=10, 3=11
    tau_sum: np.ndarray      # (nSpec, Nz+1)


def get_scattering_interface(prev: int, scatter: bool, iz: int) -> int:
    """rt_helper_functions.jl:15-33 (iz 0-based)."""
    if iz == 0:
        return 3 if scatter else 0
    if prev == 0:
        return 1 if scatter else 0
    return 3 if scatter else 2


def band_control(model: Model, m: int, ztab=None) -> BandControl:
    FT = model.FT
    props = construct_core_optical_properties(model, m, ztab=ztab, with_Z=False)
    eps = np.finfo(FT).eps
    scat, nds, itf = [], [], []
    tau_sum = np.zeros((model.nSpec, model.Nz + 1), FT)
    prev = 0
    for iz, p in enumerate(props):
        mx = np.max(p.tau * p.varpi)
        s = bool(mx > 2 * eps)
        prev = get_scattering_interface(prev, s, iz)
        scat.append(s)
        itf.append(prev)
        nds.append(get_ndoubl(mx, model.quad, FT, model.dtau_max_threshold,
                              model.dtau_min_floor) if s else 0)
        tau_sum[:, iz + 1] = tau_sum[:, iz] + p.tau
    return BandControl(scat, nds, itf, tau_sum)


# --------------------------------------------------------------------------
# elemental   src/CoreRT/CoreKernel/elemental.jl:174-462, rt_helpers.jl:15-40
# --------------------------------------------------------------------------
def expdiff_neg(a, b):
    """exp(-a) - exp(-b), rt_helpers.jl:32-40 (vectorised)."""
    a, b = np.broadcast_arrays(a, b)
    lo = np.minimum(a, b)
    d = np.abs(a - b)
    v = np.exp(-lo) * (-np.expm1(-d))
    return np.where(a == b, np.zeros_like(v), np.where(a < b, v, -v))


def _D_diag(pol_n: int, Nquad: int, FT) -> np.ndarray:
    D = np.array([1, 1, -1, -1][:pol_n], FT)
    return np.tile(D, Nquad)


def elemental(model: Model, m: int, ndoubl: int, dtau, varpi, Zpp, Zmp,
              tau_sum, F0):
    """elemental! + get_elem_rt! + get_elem_rt_SFI! + apply_D_elemental!.
    dtau, varpi, tau_sum: (nS,);  Z: (nS,N,N) or (N,N);  F0: (pol_n, nS).
    Returns r_mp, t_pp, r_pm, t_mm, j0p, j0m (r_pm/t_mm only valid when
    ndoubl == 0, as in the reference)."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    mu = q.qp_muN.astype(FT)
    eps = np.finfo(FT).eps
    wct02 = FT(0.5 if m == 0 else 0.25)
    wct = (q.wt_muN / 2 if m == 0 else q.wt_muN / 4).astype(FT)
    nS = dtau.shape[0]
    N = mu.size
    if Zpp.ndim == 2:
        Zpp = np.broadcast_to(Zpp, (nS, N, N))
        Zmp = np.broadcast_to(Zmp, (nS, N, N))
    mui = mu[None, :, None]
    muj = mu[None, None, :]
    w_j = wct[None, None, :]
    w = varpi[:, None, None]
    d = dtau[:, None, None]
    # r⁻⁺  (elemental.jl:298-302)
    r = w * Zmp * (muj / (mui + muj)) * w_j * (-np.expm1(-d * ((1 / mui) + (1 / muj))))
    # t⁺⁺ off-diagonal (elemental.jl:319-323)
    with np.errstate(divide='ignore', invalid='ignore'):
        t_off = w * Zpp * (muj / (mui - muj)) * w_j * expdiff_neg(d / mui, d / muj)
    # μ[i]==μ[j] branch (elemental.jl:304-314)
    eye = np.eye(N, dtype=bool)[None]
    same = (mui == muj)
    t_same_diag = np.exp(-d / mui) * (1 + w * Zpp * (d / mui) * wct[None, :, None])
    t_same_off = np.exp(-d / muj) * (w * Zpp * (d / mui) * w_j)
    t = np.where(same, np.where(eye, t_same_diag, t_same_off), t_off)
    # zero-weight columns (elemental.jl:325-332)
    active = (wct > eps)[None, None, :]
    r = np.where(active, r, 0).astype(FT)
    t_beer = np.where(eye, np.exp(-d / mui), 0)
    t = np.where(active, t, t_beer).astype(FT)

    # SFI (elemental.jl:348-392)
    i_start = pol_n * q.imu0
    i_end = i_start + pol_n
    ZppI = np.einsum('nik,kn->ni', Zpp[:, :, i_start:i_end], F0.astype(FT))
    ZmpI = np.einsum('nik,kn->ni', Zmp[:, :, i_start:i_end], F0.astype(FT))
    mu_i = mu[None, :]
    mu_s = mu[i_start]
    dd = dtau[:, None]
    ww = varpi[:, None]
    insun = np.zeros(N, bool)
    insun[i_start:i_end] = True
    jp_in = wct02 * ww * ZppI * (dd / mu_i) * np.exp(-dd / mu_i)
    with np.errstate(divide='ignore', invalid='ignore'):
        jp_out = wct02 * ww * ZppI * (mu_s / (mu_i - mu_s)) * expdiff_neg(dd / mu_i, dd / mu_s)
    j0p = np.where(insun[None, :], jp_in, jp_out)
    j0m = wct02 * ww * ZmpI * (mu_s / (mu_i + mu_s)) * (-np.expm1(-dd * ((1 / mu_i) + (1 / mu_s))))
    att = np.exp(-tau_sum / mu_s)[:, None]
    j0p = (j0p * att).astype(FT)
    j0m = (j0m * att).astype(FT)
    Dd = _D_diag(pol_n, q.Nquad, FT)
    if ndoubl >= 1:
        j0m = Dd[None, :] * j0m
    # apply_D_elemental! (elemental.jl:403-422)
    r_pm = t_mm = None
    if ndoubl < 1:
        sgn = (Dd[:, None] * Dd[None, :])[None]
        r_pm = sgn * r
        t_mm = sgn * t
    else:
        r = Dd[None, :, None] * r
    return r, t, r_pm, t_mm, j0p, j0m


# --------------------------------------------------------------------------
# doubling   src/CoreRT/CoreKernel/doubling.jl:38-99, rt_helpers.jl:102-166
# --------------------------------------------------------------------------
def doubling(model: Model, ndoubl: int, expk, r, t, j0p, j0m):
    """Returns r_mp, t_pp, r_pm, t_mm, j0p, j0m after ndoubl doublings and
    apply_D_matrix!/apply_D_matrix_SFI!.  Must only be called for ndoubl>=1."""
    FT = model.FT
    N = r.shape[1]
    I = np.eye(N, dtype=FT)[None]
    expk = expk.copy()
    jp = j0p[:, :, None]
    jm = j0m[:, :, None]
    for _ in range(ndoubl):
        gp = np.linalg.inv(I - r @ r)
        tt = t @ gp
        e = expk[:, None, None]
        j1p = jp * e
        j1m = jm * e
        jm_new = jm + tt @ (j1m + r @ jp)
        jp_new = j1p + tt @ (jp + r @ j1m)
        jm, jp = jm_new, jp_new
        r_new = r + (tt @ r) @ t
        t = tt @ t
        r = r_new
        expk = expk ** 2
    Dd = _D_diag(model.pol_n, model.quad.Nquad, FT)
    if model.pol_n == 1:
        return r, t, r.copy(), t.copy(), jp[:, :, 0], jm[:, :, 0]
    r = Dd[None, :, None] * r
    sgn = (Dd[:, None] * Dd[None, :])[None]
    r_pm = sgn * r
    t_mm = sgn * t
    jm = Dd[None, :, None] * jm
    return r, t, r_pm, t_mm, jp[:, :, 0], jm[:, :, 0]


# --------------------------------------------------------------------------
# interaction   src/CoreRT/CoreKernel/interaction.jl:52-266
# --------------------------------------------------------------------------
@dataclass
class Composite:
    Rmp: np.ndarray
    Rpm: np.ndarray
    Tpp: np.ndarray
    Tmm: np.ndarray
    J0p: np.ndarray      # (nS, N)
    J0m: np.ndarray


def _mv(A, x):
    return (A @ x[:, :, None])[:, :, 0]


def interaction(itf: int, C: Composite, r_mp, r_pm, t_pp, t_mm, j0p, j0m):
    FT = C.Rmp.dtype
    N = C.Rmp.shape[1]
    I = np.eye(N, dtype=FT)[None]
    if itf == 0:     # interaction.jl:52-72
        C.J0p = j0p + _mv(t_pp, C.J0p)
        C.J0m = C.J0m + _mv(C.Tmm, j0m)
        C.Tmm = t_mm @ C.Tmm
        C.Tpp = t_pp @ C.Tpp
    elif itf == 1:   # interaction.jl:96-118
        C.J0m = C.J0m + _mv(C.Tmm, _mv(r_mp, C.J0p) + j0m)
        C.J0p = j0p + _mv(t_pp, C.J0p)
        C.Rmp = (C.Tmm @ r_mp) @ C.Tpp
        C.Rpm = r_pm.copy()
        C.Tpp = t_pp @ C.Tpp
        C.Tmm = C.Tmm @ t_mm
    elif itf == 2:   # interaction.jl:141-162
        C.J0p = j0p + _mv(t_pp, C.J0p + _mv(C.Rpm, j0m))
        C.J0m = C.J0m + _mv(C.Tmm, j0m)
        C.Tpp = t_pp @ C.Tpp
        C.Tmm = C.Tmm @ t_mm
        C.Rpm = (t_pp @ C.Rpm) @ t_mm
    else:            # interaction.jl:207-266
        tmp = np.linalg.inv(I - r_mp @ C.Rpm)
        T01 = C.Tmm @ tmp
        J0m = C.J0m + _mv(T01, _mv(r_mp, C.J0p) + j0m)
        Rmp = C.Rmp + (T01 @ r_mp) @ C.Tpp
        Tmm = T01 @ t_mm
        tmp = np.linalg.inv(I - C.Rpm @ r_mp)
        T21 = t_pp @ tmp
        J0p = j0p + _mv(T21, C.J0p + _mv(C.Rpm, j0m))
        Tpp = T21 @ C.Tpp
        Rpm = r_pm + (T21 @ C.Rpm) @ t_mm
        C.J0m, C.Rmp, C.Tmm, C.J0p, C.Tpp, C.Rpm = J0m, Rmp, Tmm, J0p, Tpp, Rpm
    return C


# --------------------------------------------------------------------------
# Lambertian surface   src/CoreRT/Surfaces/lambertian_surface.jl:41-95
# --------------------------------------------------------------------------
def lambertian_surface(model: Model, m: int, tau_sum_end):
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    N = q.qp_muN.size
    nS = tau_sum_end.shape[0]
    I = np.eye(N, dtype=FT)
    t = np.broadcast_to(I, (nS, N, N))
    zeros = np.zeros((nS, N, N), FT)
    if m != 0:
        z = np.zeros((nS, N), FT)
        return zeros, zeros, t, t, z, z
    rho = FT(2) * FT(model.albedo)
    blk = np.zeros((pol_n, pol_n), FT)
    blk[0, 0] = rho
    R_surf = np.tile(blk, (q.Nquad, q.Nquad))
    I0 = np.zeros(N, FT)
    I0[pol_n * q.imu0] = 1           # pol_type.I₀ = (1,0,0,0)
    att = np.exp(-tau_sum_end / FT(q.mu0))[:, None]
    j0p = (I0[None, :] * att).astype(FT)
    j0m = (FT(q.mu0) * (R_surf @ I0)[None, :] * att).astype(FT)
    r_mp = R_surf * (q.qp_muN * q.wt_muN).astype(FT)[None, :]
    r_mp = np.broadcast_to(r_mp, (nS, N, N))
    return r_mp, zeros, t, t, j0p, j0m


# --------------------------------------------------------------------------
# Linearized doubling / interaction (config 4, LinMode): the tangent-linear rules of
# src/CoreRT/CoreKernel/doubling_lin.jl:216-335 and interaction_lin.jl:54-331.
# Derivative arrays carry a leading parameter axis P: (P, nS, N, N) / (P, nS, N).
# Pinned by central finite differences of the forward functions above
# (tests/test_oracle_lin.py), the way the reference pins its own LinMode
# (analytic vs FD, test/test_forward_lin.jl:167-244).  Not yet used by a CUDA path.
# --------------------------------------------------------------------------
def elemental_lin(model: Model, m: int, ndoubl: int, dtau, varpi, Zpp, Zmp, tau_sum, F0,
                  dtau_dot, varpi_dot, Zpp_dot, Zmp_dot, tausum_dot):
    """Per-parameter derivatives of the elemental layer: the fused chain rule of the LinMode
    elemental! (elemental_lin.jl:77-206 with lin_added_layer_all_params!,
    lin_added_layer_all_params.jl:1-136): d/dp = (d/d dtau) dtau_dot + (d/d varpi) varpi_dot +
    (d/dZ) Z_dot, and d(exp(-tau_sum/mu0))/dp for the solar sources.
    dtau_dot, varpi_dot, tausum_dot: (P,nS);  Zpp_dot, Zmp_dot: (P,nS,N,N) or (P,N,N).
    Returns (dr_mp, dt_pp, dr_pm, dt_mm, dj0p, dj0m), each with a leading P axis, with the same
    D-matrix conventions as elemental()."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    mu = q.qp_muN.astype(FT)
    eps = np.finfo(FT).eps
    wct02 = FT(0.5 if m == 0 else 0.25)
    wct = (q.wt_muN / 2 if m == 0 else q.wt_muN / 4).astype(FT)
    nS, N, P = dtau.shape[0], mu.size, dtau_dot.shape[0]
    if Zpp.ndim == 2:
        Zpp = np.broadcast_to(Zpp, (nS, N, N)); Zmp = np.broadcast_to(Zmp, (nS, N, N))
    if Zpp_dot.ndim == 3:
        Zpp_dot = np.broadcast_to(Zpp_dot[:, None], (P, nS, N, N)); Zmp_dot = np.broadcast_to(Zmp_dot[:, None], (P, nS, N, N))
    mui, muj = mu[None, :, None], mu[None, None, :]
    w_j = wct[None, None, :]
    w, d = varpi[:, None, None], dtau[:, None, None]
    wd, dd_ = varpi_dot[:, :, None, None], dtau_dot[:, :, None, None]
    c = (1 / mui) + (1 / muj)
    g1 = (muj / (mui + muj)) * w_j
    E1, dE1 = -np.expm1(-d * c), c * np.exp(-d * c)
    # r-+ = w Zmp g1 E1(d)
    dr = (wd * Zmp[None] + w[None] * Zmp_dot) * (g1 * E1)[None] + (w * Zmp * g1 * dE1)[None] * dd_
    with np.errstate(divide='ignore', invalid='ignore'):
        g2 = (muj / (mui - muj)) * w_j
        E2 = expdiff_neg(d / mui, d / muj)
        dE2 = -np.exp(-d / mui) / mui + np.exp(-d / muj) / muj
        dt_off = (wd * Zpp[None] + w[None] * Zpp_dot) * (g2 * E2)[None] + (w * Zpp * g2 * dE2)[None] * dd_
    eye = np.eye(N, dtype=bool)[None]
    same = (mui == muj)
    wi = wct[None, :, None]
    ei, ej = np.exp(-d / mui), np.exp(-d / muj)
    # exp(-d/mu_i) (1 + w Zpp (d/mu_i) wct_i)
    dt_sd = (ei * (d / mui) * wi)[None] * (wd * Zpp[None] + w[None] * Zpp_dot) \
        + (-(1 / mui) * ei * (1 + w * Zpp * (d / mui) * wi) + ei * w * Zpp * wi / mui)[None] * dd_
    # exp(-d/mu_j) w Zpp (d/mu_i) w_j
    dt_so = (ej * (d / mui) * w_j)[None] * (wd * Zpp[None] + w[None] * Zpp_dot) \
        + (-(1 / muj) * ej * w * Zpp * (d / mui) * w_j + ej * w * Zpp * w_j / mui)[None] * dd_
    dt = np.where(same[None], np.where(eye[None], dt_sd, dt_so), dt_off)
    active = (wct > eps)[None, None, None, :]
    dr = np.where(active, dr, 0).astype(FT)
    dt_beer = np.where(eye[None], (-ei / mui)[None] * dd_, 0)
    dt = np.where(active, dt, dt_beer).astype(FT)
    # solar sources
    i_start = pol_n * q.imu0
    i_end = i_start + pol_n
    F = F0.astype(FT)
    ZppI = np.einsum('nik,kn->ni', Zpp[:, :, i_start:i_end], F)
    ZmpI = np.einsum('nik,kn->ni', Zmp[:, :, i_start:i_end], F)
    dZppI = np.einsum('pnik,kn->pni', Zpp_dot[:, :, :, i_start:i_end], F)
    dZmpI = np.einsum('pnik,kn->pni', Zmp_dot[:, :, :, i_start:i_end], F)
    mu_i, mu_s = mu[None, :], mu[i_start]
    dd, ww = dtau[:, None], varpi[:, None]
    ddp, wwp = dtau_dot[:, :, None], varpi_dot[:, :, None]
    insun = np.zeros(N, bool); insun[i_start:i_end] = True
    e_i = np.exp(-dd / mu_i)
    f_in, df_in = (dd / mu_i) * e_i, e_i / mu_i - (dd / mu_i) * e_i / mu_i
    with np.errstate(divide='ignore', invalid='ignore'):
        gs2 = mu_s / (mu_i - mu_s)
        f_out = gs2 * expdiff_neg(dd / mu_i, dd / mu_s)
        df_out = gs2 * (-np.exp(-dd / mu_i) / mu_i + np.exp(-dd / mu_s) / mu_s)
    fp = np.where(insun[None, :], f_in, f_out)
    dfp = np.where(insun[None, :], df_in, df_out)
    cs = (1 / mu_i) + (1 / mu_s)
    fm = (mu_s / (mu_i + mu_s)) * (-np.expm1(-dd * cs))
    dfm = (mu_s / (mu_i + mu_s)) * cs * np.exp(-dd * cs)
    att = np.exp(-tau_sum / mu_s)[:, None]
    datt = -(att / mu_s)[None] * tausum_dot[:, :, None]
    jp0 = wct02 * ww * ZppI * fp
    jm0 = wct02 * ww * ZmpI * fm
    djp0 = wct02 * ((wwp * ZppI[None] + ww[None] * dZppI) * fp[None] + (ww * ZppI * dfp)[None] * ddp)
    djm0 = wct02 * ((wwp * ZmpI[None] + ww[None] * dZmpI) * fm[None] + (ww * ZmpI * dfm)[None] * ddp)
    dj0p = (djp0 * att[None] + jp0[None] * datt).astype(FT)
    dj0m = (djm0 * att[None] + jm0[None] * datt).astype(FT)
    Dd = _D_diag(pol_n, q.Nquad, FT)
    if ndoubl >= 1:
        dj0m = Dd[None, None, :] * dj0m
    dr_pm = dt_mm = None
    if ndoubl < 1:
        sgn = (Dd[:, None] * Dd[None, :])[None, None]
        dr_pm, dt_mm = sgn * dr, sgn * dt
    else:
        dr = Dd[None, None, :, None] * dr
    return dr, dt, dr_pm, dt_mm, dj0p, dj0m


def _mvp(A, x):
    """A (..., N, N) times x (..., N), batched with broadcasting over leading axes."""
    return np.einsum('...ij,...j->...i', A, x)


def doubling_lin(model: Model, ndoubl: int, expk, r, t, j0p, j0m, dr, dt, dj0p, dj0m, dtau_dot):
    """doubling_allparams_helper! (doubling_lin.jl:216-335).  Forward inputs as doubling();
    dr, dt (P,nS,N,N), dj0p, dj0m (P,nS,N): per-parameter derivatives of the elemental
    layer; dtau_dot (P,nS): d(dtau)/dp, which drives d(expk)/dp = -expk/mu0 * dtau_dot (:257-259).
    Returns (forward tuple as doubling(), derivative tuple (dr_mp, dt_pp, dr_pm, dt_mm, djp, djm))."""
    FT = model.FT
    N = r.shape[1]
    I = np.eye(N, dtype=FT)[None]
    expk = expk.copy()
    dexpk = -(expk / FT(model.quad.mu0))[None, :] * dtau_dot
    r, t, jp, jm = r.copy(), t.copy(), j0p.copy(), j0m.copy()
    dr, dt, djp, djm = dr.copy(), dt.copy(), dj0p.copy(), dj0m.copy()
    for _ in range(ndoubl):
        gp = np.linalg.inv(I - r @ r)                                   # :268
        tt = t @ gp
        dgp = gp[None] @ (dr @ r[None] + r[None] @ dr) @ gp[None]        # :273
        dtt = dt @ gp[None] + t[None] @ dgp                             # :274
        e = expk[:, None]
        j1p, j1m = jp * e, jm * e                                       # :283-284
        a_m = j1m + _mvp(r, jp)                                         # :287
        a_p = jp + _mvp(r, j1m)                                         # :288
        dj1p = djp * e[None] + jp[None] * dexpk[:, :, None]             # :291
        dj1m = djm * e[None] + jm[None] * dexpk[:, :, None]             # :292
        dexpk = 2 * expk[None, :] * dexpk                               # :294
        djm_new = djm + _mvp(dtt, a_m[None]) + _mvp(tt[None], dj1m + _mvp(dr, jp[None]) + _mvp(r[None], djp))   # :296-298
        djp_new = dj1p + _mvp(dtt, a_p[None]) + _mvp(tt[None], djp + _mvp(dr, j1m[None]) + _mvp(r[None], dj1m))  # :299-301
        djm, djp = djm_new, djp_new
        jm, jp = jm + _mvp(tt, a_m), j1p + _mvp(tt, a_p)                # :305-306
        expk = expk ** 2
        rt = r @ t                                                      # :310
        dr_new = dr + dtt @ rt[None] + tt[None] @ (dr @ t[None] + r[None] @ dt)   # :314-317
        dt_new = dtt @ t[None] + tt[None] @ dt                          # :318-319
        dr, dt = dr_new, dt_new
        r, t = r + tt @ rt, tt @ t                                      # :323-324
    Dd = _D_diag(model.pol_n, model.quad.Nquad, FT)
    if model.pol_n == 1:
        return (r, t, r.copy(), t.copy(), jp, jm), (dr, dt, dr.copy(), dt.copy(), djp, djm)
    sgn = (Dd[:, None] * Dd[None, :])
    r = Dd[None, :, None] * r                                           # apply_D_matrix! (:330-333)
    dr = Dd[None, None, :, None] * dr
    jm = Dd[None, :] * jm                                               # apply_D_matrix_SFI! (:334)
    djm = Dd[None, None, :] * djm
    return (r, t, sgn[None] * r, sgn[None] * t, jp, jm), (dr, dt, sgn[None, None] * dr, sgn[None, None] * dt, djp, djm)


def interaction_lin(itf: int, C: Composite, dC: Composite, r_mp, r_pm, t_pp, t_mm, j0p, j0m,
                    dr_mp, dr_pm, dt_pp, dt_mm, dj0p, dj0m):
    """interaction_helper!(::ScatteringInterface_xx, ...) of LinMode (interaction_lin.jl:54-331).
    C / dC: composite layer and its per-parameter derivatives (fields (P,nS,...)); both are
    updated and returned."""
    FT = C.Rmp.dtype
    N = C.Rmp.shape[1]
    I = np.eye(N, dtype=FT)[None]
    b = lambda x: x[None]
    if itf == 0:     # :54-87
        dJ0p = dj0p + _mvp(b(t_pp), dC.J0p) + _mvp(dt_pp, b(C.J0p))
        dJ0m = dC.J0m + _mvp(b(C.Tmm), dj0m) + _mvp(dC.Tmm, b(j0m))
        J0p = j0p + _mvp(t_pp, C.J0p)
        J0m = C.J0m + _mvp(C.Tmm, j0m)
        dTmm = dt_mm @ b(C.Tmm) + b(t_mm) @ dC.Tmm
        dTpp = dt_pp @ b(C.Tpp) + b(t_pp) @ dC.Tpp
        C.Tmm, C.Tpp = t_mm @ C.Tmm, t_pp @ C.Tpp
        C.J0p, C.J0m, dC.J0p, dC.J0m, dC.Tmm, dC.Tpp = J0p, J0m, dJ0p, dJ0m, dTmm, dTpp
    elif itf == 1:   # :96-141
        x = _mvp(r_mp, C.J0p) + j0m
        dJ0m = dC.J0m + _mvp(dC.Tmm, b(x)) + _mvp(b(C.Tmm), _mvp(dr_mp, b(C.J0p)) + _mvp(b(r_mp), dC.J0p) + dj0m)
        dJ0p = dj0p + _mvp(dt_pp, b(C.J0p)) + _mvp(b(t_pp), dC.J0p)
        J0m = C.J0m + _mvp(C.Tmm, x)
        J0p = j0p + _mvp(t_pp, C.J0p)
        dRmp = dC.Tmm @ b(r_mp) @ b(C.Tpp) + b(C.Tmm) @ dr_mp @ b(C.Tpp) + b(C.Tmm) @ b(r_mp) @ dC.Tpp
        dTpp = dt_pp @ b(C.Tpp) + b(t_pp) @ dC.Tpp
        dTmm = dC.Tmm @ b(t_mm) + b(C.Tmm) @ dt_mm
        C.Rmp = (C.Tmm @ r_mp) @ C.Tpp
        C.Rpm = r_pm.copy()
        C.Tpp, C.Tmm = t_pp @ C.Tpp, C.Tmm @ t_mm
        C.J0p, C.J0m = J0p, J0m
        dC.Rmp, dC.Rpm, dC.Tpp, dC.Tmm, dC.J0p, dC.J0m = dRmp, dr_pm.copy(), dTpp, dTmm, dJ0p, dJ0m
    elif itf == 2:   # :150-196
        y = C.J0p + _mvp(C.Rpm, j0m)
        dJ0p = dj0p + _mvp(dt_pp, b(y)) + _mvp(b(t_pp), dC.J0p + _mvp(dC.Rpm, b(j0m)) + _mvp(b(C.Rpm), dj0m))
        dJ0m = dC.J0m + _mvp(dC.Tmm, b(j0m)) + _mvp(b(C.Tmm), dj0m)
        J0p = j0p + _mvp(t_pp, y)
        J0m = C.J0m + _mvp(C.Tmm, j0m)
        dTpp = dt_pp @ b(C.Tpp) + b(t_pp) @ dC.Tpp
        dTmm = dC.Tmm @ b(t_mm) + b(C.Tmm) @ dt_mm
        dRpm = dt_pp @ b(C.Rpm) @ b(t_mm) + b(t_pp) @ dC.Rpm @ b(t_mm) + b(t_pp) @ b(C.Rpm) @ dt_mm
        C.Tpp, C.Tmm = t_pp @ C.Tpp, C.Tmm @ t_mm
        C.Rpm = (t_pp @ C.Rpm) @ t_mm
        C.J0p, C.J0m = J0p, J0m
        dC.Tpp, dC.Tmm, dC.Rpm, dC.J0p, dC.J0m = dTpp, dTmm, dRpm, dJ0p, dJ0m
    else:            # :217-331
        g1 = np.linalg.inv(I - r_mp @ C.Rpm)
        T01 = C.Tmm @ g1
        rT = r_mp @ C.Tpp
        dg1 = b(g1) @ (dr_mp @ b(C.Rpm) + b(r_mp) @ dC.Rpm) @ b(g1)                      # :247
        dT01 = dC.Tmm @ b(g1) + b(C.Tmm) @ dg1                                           # :248
        dRmp = dC.Rmp + dT01 @ b(rT) + b(T01) @ (dr_mp @ b(C.Tpp) + b(r_mp) @ dC.Tpp)    # :250-253
        dTmm = dT01 @ b(t_mm) + b(T01) @ dt_mm                                           # :256
        x = _mvp(r_mp, C.J0p) + j0m
        dJ0m = dC.J0m + _mvp(dT01, b(x)) + _mvp(b(T01), _mvp(dr_mp, b(C.J0p)) + _mvp(b(r_mp), dC.J0p) + dj0m)   # :265-267
        J0m = C.J0m + _mvp(T01, x)
        Rmp = C.Rmp + T01 @ rT
        Tmm = T01 @ t_mm
        g2 = np.linalg.inv(I - C.Rpm @ r_mp)
        T21 = t_pp @ g2
        Rt = C.Rpm @ t_mm
        dg2 = b(g2) @ (b(C.Rpm) @ dr_mp + dC.Rpm @ b(r_mp)) @ b(g2)                      # :288
        dT21 = dt_pp @ b(g2) + b(t_pp) @ dg2                                             # :289
        dTpp = dT21 @ b(C.Tpp) + b(T21) @ dC.Tpp                                         # :292
        dRpm = dr_pm + dT21 @ b(Rt) + b(T21) @ (dC.Rpm @ b(t_mm) + b(C.Rpm) @ dt_mm)     # :295-296
        y = C.J0p + _mvp(C.Rpm, j0m)
        dJ0p = dj0p + _mvp(dT21, b(y)) + _mvp(b(T21), dC.J0p + _mvp(dC.Rpm, b(j0m)) + _mvp(b(C.Rpm), dj0m))      # :301-306
        J0p = j0p + _mvp(T21, y)
        Tpp = T21 @ C.Tpp
        Rpm = r_pm + T21 @ Rt
        C.J0m, C.Rmp, C.Tmm, C.J0p, C.Tpp, C.Rpm = J0m, Rmp, Tmm, J0p, Tpp, Rpm
        dC.J0m, dC.Rmp, dC.Tmm, dC.J0p, dC.Tpp, dC.Rpm = dJ0m, dRmp, dTmm, dJ0p, dTpp, dRpm
    return C, dC


# --------------------------------------------------------------------------
# BRDF surfaces   src/CoreRT/Surfaces/rpv_surface.jl:51-97 (layer), :113-165 (RPV)
# --------------------------------------------------------------------------
def brdf_surface(model: Model, m: int, tau_sum_end):
    """create_surface_layer!(::AbstractSurfaceType, ...) (rpv_surface.jl:51-97) with the
    Fourier moment R_surf = model.surface_rho[m] (already carrying the factor 2 of m == 0,
    :69-73).  Unlike the Lambertian case every moment has a reflecting surface."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    N = q.qp_muN.size
    nS = tau_sum_end.shape[0]
    I = np.eye(N, dtype=FT)
    t = np.broadcast_to(I, (nS, N, N))
    zeros = np.zeros((nS, N, N), FT)
    R_surf = np.asarray(model.surface_rho[m], FT)
    I0 = np.zeros(N, FT)
    I0[pol_n * q.imu0] = 1
    att = np.exp(-tau_sum_end / FT(q.mu0))[:, None]
    j0p = (I0[None, :] * att).astype(FT)
    j0m = (FT(q.mu0) * (R_surf @ I0)[None, :] * att).astype(FT)
    r_mp = R_surf * (q.qp_muN * q.wt_muN).astype(FT)[None, :]
    r_mp = np.broadcast_to(r_mp, (nS, N, N))
    return r_mp, zeros, t, t, j0p, j0m


def rpv_reflectance(rho0, rho_c, k, theta_hs, mu_i, mu_r, dphi):
    """Scalar RPV BRDF (Rahman, Pinty, Verstraete 1993) with the reference's sign
    conventions, rpv_surface.jl:113-144."""
    th_i = math.acos(mu_i)
    th_r = math.acos(mu_r)
    cosg = -mu_i * mu_r + math.sin(th_i) * math.sin(th_r) * math.cos(dphi)
    G = math.sqrt(math.tan(th_i) ** 2 + math.tan(th_r) ** 2 + 2 * math.tan(th_i) * math.tan(th_r) * math.cos(dphi))
    M = (mu_i * mu_r) ** (k - 1) / (mu_i + mu_r) ** (1 - k)
    th = -theta_hs
    F = (1 - th * th) / (1 + th * th + 2 * th * cosg) ** 1.5
    H = 1 + (1 - rho_c) / (1 + G)
    return rho0 * M * F * H


def rossli_reflectance(fiso, fvol, fgeo, mu_i, mu_r, dphi):
    """Scalar Ross-Li BRDF (rossli_surface.jl: reflectance + RossLi_K_vol + RossLi_K_geo, RAMI defaults
    h/b = 2, b/r = 1)."""
    dphi = math.pi - dphi
    th_i, th_r = math.acos(mu_i), math.acos(mu_r)
    xi = math.acos(max(-1.0, min(1.0, math.cos(th_i) * math.cos(th_r) + math.sin(th_i) * math.sin(th_r) * math.cos(dphi))))
    k_vol = ((math.pi / 2 - xi) * math.cos(xi) + math.sin(xi)) / (math.cos(th_i) + math.cos(th_r)) - math.pi / 4
    tip, trp = math.atan(math.tan(th_i) * 1.0), math.atan(math.tan(th_r) * 1.0)
    xip = math.acos(max(-1.0, min(1.0, math.cos(tip) * math.cos(trp) + math.sin(tip) * math.sin(trp) * math.cos(dphi))))
    D = math.sqrt(max(0.0, math.tan(tip) ** 2 + math.tan(trp) ** 2 - 2 * math.tan(tip) * math.tan(trp) * math.cos(dphi)))
    sec = 1 / math.cos(tip) + 1 / math.cos(trp)
    ct = 2.0 * math.sqrt(D * D + (math.tan(tip) * math.tan(trp) * math.sin(dphi)) ** 2) / sec
    ct = max(-1.0, min(1.0, ct))
    t = math.acos(ct)
    O = (1 / math.pi) * (t - math.sin(t) * math.cos(t)) * sec
    k_geo = O - sec + 0.5 * (1 + math.cos(xip)) / (math.cos(tip) * math.cos(trp))
    return fiso + fvol * k_vol + fgeo * k_geo


def brdf_surface_rho(quad: QuadPoints, m_max: int, fn, nquad: int = 100):
    """reflectance(brdf, pol_type, mu, m) (rpv_surface.jl:166-196) for a scalar BRDF fn(mu_i, mu_r, dphi), with
    the factor 2 of create_surface_layer! (:69-73).  Pure-Python loops: small cases only."""
    pol_n = quad.pol_n
    mu = np.asarray(quad.qp_mu, float)
    nq = mu.size
    phi, w = gauleg(nquad, 0.0, math.pi)
    out = np.zeros((m_max + 1, nq * pol_n, nq * pol_n))
    for i in range(nq):
        for j in range(nq):
            vals = [fn(mu[i], mu[j], x) for x in phi]
            for m in range(m_max + 1):
                c = sum(wx * v * math.cos(m * x) for x, wx, v in zip(phi, w, vals))
                out[m, i * pol_n, j * pol_n] = 2.0 * c / math.pi
    return out


def rpv_surface_rho(quad: QuadPoints, m_max: int, rho0, rho_c, k, theta_hs, nquad: int = 100):
    """Fourier moments of the RPV reflectance matrix, reflectance(brdf, pol_type, mu, m)
    (rpv_surface.jl:166-196: 100-point Gauss-Legendre in azimuth over [0, pi], only the
    Stokes-I block is non-zero) times the factor 2 that create_surface_layer! applies to
    m == 0 (:69-73; for m > 0 the 2 is `ff`).  Pure-Python loops: small cases only."""
    pol_n = quad.pol_n
    mu = np.asarray(quad.qp_mu, float)
    nq = mu.size
    phi, w = gauleg(nquad, 0.0, math.pi)
    out = np.zeros((m_max + 1, nq * pol_n, nq * pol_n))
    for m in range(m_max + 1):
        for i in range(nq):
            for j in range(nq):
                c = 0.0
                for x, wx in zip(phi, w):
                    c += wx * rpv_reflectance(rho0, rho_c, k, theta_hs, mu[i], mu[j], x) * math.cos(m * x)
                out[m, i * pol_n, j * pol_n] = 2.0 * c / math.pi
    return out


# --------------------------------------------------------------------------
# Cox-Munk ocean surface   src/CoreRT/Surfaces/coxmunk_surface.jl, fresnel.jl
# (scalar restatement, pure-Python loops: small quadratures only)
# --------------------------------------------------------------------------
# water_refractive_index(550 nm) (water_refraction.jl: Segelstein 1981 table row 550 nm).  Every reference call
# site evaluates the default `n_water = _get_n_water(surf, FT(550))` (coxmunk_surface.jl:262,465,514).
N_WATER_550 = complex(1.328, 5.69e-9)


@dataclass
class CoxMunk:
    """CoxMunkSurface (types.jl:525-536)."""
    wind_speed: float
    n_water: complex = N_WATER_550
    whitecap_albedo: float = 0.22
    include_whitecaps: bool = True
    shadowing: bool = True


def wind_to_sigma2(U):                                   # coxmunk_surface.jl:22
    return 0.003 + 0.00512 * U


def cox_munk_pdf(zx, zy, s2):                            # :25-27
    return math.exp(-(zx * zx + zy * zy) / (2 * s2)) / (2 * math.pi * s2)


def whitecap_fraction(U):                                # :30-33
    return 0.0 if U <= 0 else 2.95e-6 * U ** 3.52


def smith_lambda(mu, s2):                                # :41-50
    if mu <= 0:
        return 1e10
    sig = math.sqrt(s2)
    cot = mu / math.sqrt(max(1e-30, 1.0 - mu * mu))
    nu = cot / (math.sqrt(2.0) * sig)
    lam = (math.exp(-nu * nu) / (math.sqrt(2 * math.pi) * nu) - math.erfc(nu)) / 2
    return max(0.0, lam)


def shadow_factor(mu_i, mu_r, s2):                       # :53-55
    return 1.0 / (1.0 + smith_lambda(mu_i, s2) + smith_lambda(mu_r, s2))


def fresnel_coefficients(n_rel: complex, cos_i: float):  # fresnel.jl:25-33
    import cmath
    sin2 = max(0.0, 1.0 - cos_i * cos_i)
    cos_t = cmath.sqrt(1.0 - sin2 / (n_rel * n_rel))
    r_s = (cos_i - n_rel * cos_t) / (cos_i + n_rel * cos_t)
    r_p = (n_rel * cos_i - cos_t) / (n_rel * cos_i + cos_t)
    return r_s, r_p


def fresnel_mueller(r_s: complex, r_p: complex, n: int) -> np.ndarray:   # fresnel.jl:50-88
    rs2, rp2 = abs(r_s) ** 2, abs(r_p) ** 2
    rsp = r_s * r_p.conjugate()
    M = np.zeros((4, 4))
    M[0, 0] = M[1, 1] = (rs2 + rp2) / 2
    M[0, 1] = M[1, 0] = (rs2 - rp2) / 2
    M[2, 2] = M[3, 3] = rsp.real
    M[2, 3] = rsp.imag
    M[3, 2] = -rsp.imag
    return M[:n, :n]


def stokes_rotation_matrix(phi: float, n: int) -> np.ndarray:            # fresnel.jl:104-132
    c2, s2 = math.cos(2 * phi), math.sin(2 * phi)
    L = np.eye(4)
    L[1, 1] = c2; L[1, 2] = s2
    L[2, 1] = -s2; L[2, 2] = c2
    return L[:n, :n]


def coxmunk_geometry(mu_i, mu_r, dphi):                  # coxmunk_surface.jl:124-250
    sin_i = math.sqrt(max(0.0, 1 - mu_i * mu_i))
    sin_r = math.sqrt(max(0.0, 1 - mu_r * mu_r))
    cd, sd = math.cos(dphi), math.sin(dphi)
    nx = -sin_i + sin_r * cd
    ny = sin_r * sd
    nz = mu_i + mu_r
    nn = math.sqrt(nx * nx + ny * ny + nz * nz)
    if nn < 1e-15:
        return 1.0, 1.0, 0.0, 0.0, 0.0, 0.0
    nx, ny, nz = nx / nn, ny / nn, nz / nn
    cos_b = max(1e-10, nz)
    cos_loc = min(max((mu_i + mu_r) / (2 * cos_b), 0.0), 1.0)
    zx, zy = -nx / cos_b, -ny / cos_b
    cosT = -mu_i * mu_r + sin_i * sin_r * cd
    sinT = math.sqrt(max(0.0, 1 - cosT * cosT))
    a1 = a2 = 0.0
    if sinT >= 1e-12:
        sp = np.array([-mu_i * sin_r * sd, mu_i * sin_r * cd - sin_i * mu_r, -sin_i * sin_r * sd])
        ip = np.array([mu_i * ny, -mu_i * nx - sin_i * nz, sin_i * ny])
        msp, mip = math.sqrt(sp @ sp), math.sqrt(ip @ ip)
        if not (msp < 1e-15 or mip < 1e-15):
            c = min(max((sp @ ip) / (msp * mip), -1.0), 1.0)
            cr = np.cross(sp, ip)
            sg = sin_i * cr[0] + (-mu_i) * cr[2]
            a1 = math.acos(c) if sg >= 0 else -math.acos(c)
        rp = np.array([(-sin_r * sd) * nz - mu_r * ny, mu_r * nx - (-sin_r * cd) * nz,
                       (-sin_r * cd) * ny - (-sin_r * sd) * nx])
        mrp = math.sqrt(rp @ rp)
        if not (msp < 1e-15 or mrp < 1e-15):
            c = min(max((sp @ rp) / (msp * mrp), -1.0), 1.0)
            cr = np.cross(sp, rp)
            sg = (-sin_r * cd) * cr[0] + (-sin_r * sd) * cr[1] + mu_r * cr[2]
            a2 = math.acos(c) if sg >= 0 else -math.acos(c)
    return cos_b, cos_loc, zx, zy, a1, a2


def coxmunk_brdf_mueller(surf: CoxMunk, n: int, mu_i, mu_r, dphi) -> np.ndarray:   # :259-301
    """(first argument pair as the reference's: its callers pass (reflected, incident) = (mu[i], mu[j]))"""
    s2 = wind_to_sigma2(surf.wind_speed)
    cos_b, cos_loc, zx, zy, a1, a2 = coxmunk_geometry(mu_i, mu_r, dphi)
    P = cox_munk_pdf(zx, zy, s2)
    r_s, r_p = fresnel_coefficients(surf.n_water, cos_loc)
    M = stokes_rotation_matrix(a2, n) @ fresnel_mueller(r_s, r_p, n) @ stokes_rotation_matrix(-a1, n)
    pre = P / (4 * mu_i * mu_r * cos_b ** 4)
    if surf.shadowing:
        pre *= shadow_factor(mu_i, mu_r, s2)
    rho = pre * M
    if surf.include_whitecaps:
        f = whitecap_fraction(surf.wind_speed)
        wc = np.zeros((n, n))
        wc[0, 0] = surf.whitecap_albedo / math.pi
        return (1 - f) * rho + f * wc
    return rho


def _azimuthal_kernel(si: int, sj: int, m: int, dphi: float) -> float:    # :444-452 (0-based si, sj)
    return math.cos(m * dphi) if ((si <= 1) == (sj <= 1)) else math.sin(m * dphi)


def coxmunk_surface_rho(quad: QuadPoints, m_max: int, surf: CoxMunk, nquad: int = 100) -> np.ndarray:
    """reflectance(surf::CoxMunkSurface, pol_type, mu, m) (:464-498) for m = 0..m_max with the factor 2 of
    create_surface_layer! at m = 0 (rpv_surface.jl:69-73; `ff` supplies it for m > 0): (m_max+1, N, N)."""
    n = quad.pol_n
    mu = np.asarray(quad.qp_mu, float)
    nq = mu.size
    phi, w = gauleg(nquad, 0.0, math.pi)
    out = np.zeros((m_max + 1, nq * n, nq * n))
    for x, wx in zip(phi, w):
        for j in range(nq):
            for i in range(nq):
                M = coxmunk_brdf_mueller(surf, n, mu[i], mu[j], x)
                for m in range(m_max + 1):
                    cm, sm = math.cos(m * x), math.sin(m * x)
                    for sj in range(n):
                        for si in range(n):
                            az = cm if ((si <= 1) == (sj <= 1)) else sm
                            out[m, i * n + si, j * n + sj] += wx * M[si, sj] * az
    return 2.0 * out / math.pi


def coxmunk_ss_correction(surf: CoxMunk, pol_n: int, vza, vaz, mu0: float, m_max: int, nquad: int = 100) -> np.ndarray:
    """The geometry factor of apply_ss_correction! (:509-555): (nVZA, pol_n) array of
    M_exact[si,1] - M_fourier[si,1]; R_SFI[iv,si,s] += mu0 exp(-tau_total[s]/mu0) * that."""
    phi, w = gauleg(nquad, 0.0, math.pi)
    out = np.zeros((len(vza), pol_n))
    for iv in range(len(vza)):
        mu_v = cosd(vza[iv])
        dphi = math.radians(vaz[iv])
        Mex = coxmunk_brdf_mueller(surf, pol_n, mu_v, mu0, dphi)
        Ms = [coxmunk_brdf_mueller(surf, pol_n, mu_v, mu0, x) for x in phi]
        for si in range(pol_n):
            four = 0.0
            for m in range(m_max + 1):
                wm = 0.5 if m == 0 else 1.0
                ff = 1.0 if m == 0 else 2.0
                coeff = ff * sum(wx * M[si, 0] * _azimuthal_kernel(si, 0, m, x) for x, wx, M in zip(phi, w, Ms)) / math.pi
                four += wm * _azimuthal_kernel(si, 0, m, dphi) * coeff
            out[iv, si] = Mex[si, 0] - four
    return out


# --------------------------------------------------------------------------
# rt_run (noRS, SFI)   src/CoreRT/rt_run.jl:238-539, rt_kernel.jl:175-250
# --------------------------------------------------------------------------
def thermal_elemental(model: Model, dtau, varpi, B):
    """contribute!(::PreparedThermalEmission, ...) (Sources/thermal_emission.jl:236-316) at m = 0: the Planck volume
    source of one elemental layer, 2 pi (1 - varpi) B (1 - exp(-dtau/mu_i)) on the Stokes-I rows of j0+ and j0-."""
    FT = model.FT
    q = model.quad
    mu = q.qp_muN.astype(FT)
    j = np.zeros((dtau.shape[0], mu.size), FT)
    rows = np.arange(0, mu.size, model.pol_n)
    ok = mu[rows] > np.finfo(FT).eps
    coeff = (FT(2 * math.pi) * (1 - varpi) * B)[:, None]
    j[:, rows[ok]] = coeff * (-np.expm1(-dtau[:, None] / mu[rows[ok]][None, :]))
    return j


def rt_kernel_layer(model: Model, m: int, iz: int, ctl: BandControl, props: CoreProps,
                    C: Optional[Composite], F0, tau_sum_iz, thermal_B=None):
    """One layer of rt_kernel!(::noRS, ...) for the selected spectral points.
    Returns the updated composite (created at iz == 0).  thermal_B: B_layer[iz] of the selected points, or None."""
    FT = model.FT
    q = model.quad
    nS = props.tau.shape[0]
    N = q.qp_muN.size
    if ctl.scatter[iz]:
        nd = ctl.ndoubl[iz]
        dtau = (props.tau / FT(2 ** nd)).astype(FT)
        expk = np.exp(-dtau / FT(q.mu0))
        r, t, r_pm, t_mm, j0p, j0m = elemental(model, m, nd, dtau, props.varpi,
                                               props.Zpp, props.Zmp, tau_sum_iz, F0)
        if thermal_B is not None and m == 0:
            # the :thermal slot of j0_by_src (rt_kernel.jl:211-225): its own doubling with expk = ones
            # (doubling.jl:66-80, apply_D on the slot :92-94); the interaction step then updates every slot with the
            # same operators (interaction.jl:231-259) and the outputs add (postprocessing_vza.jl:67-80), so the doubled
            # slot is joined with the solar vectors here
            jt = thermal_elemental(model, dtau, props.varpi, thermal_B)
            jtp, jtm = jt, jt.copy()
            if nd > 0:
                _, _, _, _, jtp, jtm = doubling(model, nd, np.ones_like(expk), r, t, jtp, jtm)
        else:
            jtp = jtm = None
        if nd > 0:
            r, t, r_pm, t_mm, j0p, j0m = doubling(model, nd, expk, r, t, j0p, j0m)
        if jtp is not None:
            j0p, j0m = j0p + jtp, j0m + jtm
    else:
        # zero_added_noscat!, rt_helpers.jl:174-180 (j₀⁺ is left untouched by the
        # reference; it is never read by the 00/10 interaction with zero sources
        # except through j0p, which for a non-scattering first layer is zero).
        r = np.zeros((nS, N, N), FT)
        r_pm = np.zeros((nS, N, N), FT)
        tb = np.exp(-props.tau[:, None] / q.qp_muN.astype(FT)[None, :])
        t = np.zeros((nS, N, N), FT)
        idx = np.arange(N)
        t[:, idx, idx] = tb
        t_mm = t.copy()
        j0p = np.zeros((nS, N), FT)
        j0m = np.zeros((nS, N), FT)
    if iz == 0:
        return Composite(r.copy(), r_pm.copy(), t.copy(), t_mm.copy(), j0p.copy(), j0m.copy())
    return interaction(ctl.interface[iz], C, r, r_pm, t, t_mm, j0p, j0m)


@dataclass
class Prepared:
    """m-independent control flow + per-m Z tables, shareable between the
    spectral chunks of one band (used by the threaded CPU baseline)."""
    ctl: BandControl
    ztabs: list


def prepare(model: Model) -> Prepared:
    ztabs = [z_tables(model, m) for m in range(model.m_max + 1)]
    return Prepared(band_control(model, 0, ztabs[0]), ztabs)


def rt_run(model: Model, sel=slice(None), return_streams: bool = False, prep: Optional[Prepared] = None):
    """rt_run(::noRS, model, iBand) for one band, SFI = true.  ``sel`` picks a
    subset of spectral points; control flow (ndoubl, scatter, interface) is
    always derived from the full band."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    nS_all = model.nSpec
    idx = np.arange(nS_all)[sel]
    nS = idx.size
    nV = len(model.vza)
    R_SFI = np.zeros((nV, pol_n, nS), FT)
    T_SFI = np.zeros((nV, pol_n, nS), FT)
    hdr = np.zeros((nV, pol_n, nS), FT)
    bhr_uw = np.zeros((pol_n, nS), FT)
    bhr_dw = np.zeros((pol_n, nS), FT)
    F0 = model.F0
    if F0 is None:
        F0 = np.zeros((pol_n, nS_all), FT)
        F0[0, :] = 1
    F0 = F0[:, idx]
    streams = []
    for m in range(model.m_max + 1):
        weight = FT(0.5 / math.pi) if m == 0 else FT(1.0 / math.pi)
        if prep is not None:
            ctl = prep.ctl
            props = construct_core_optical_properties(model, m, idx, prep.ztabs[m])
        else:
            ctl = band_control(model, m)
            props = construct_core_optical_properties(model, m, idx)
        C = None
        for iz in range(model.Nz):
            C = rt_kernel_layer(model, m, iz, ctl, props[iz], C, F0,
                                ctl.tau_sum[idx, iz],
                                None if model.thermal_B is None else np.asarray(model.thermal_B, FT)[idx, iz])
        # surface + final interaction (rt_run.jl:423-465)
        if model.surface_rho is not None:
            s = brdf_surface(model, m, ctl.tau_sum[idx, -1])
        else:
            s = lambertian_surface(model, m, ctl.tau_sum[idx, -1])
        r_mp_s, r_pm_s, t_s, t_mm_s, j0p_s, j0m_s = s
        if model.sif0 is not None and model.surface_rho is None and m == 0:
            # surface_source_contribute!(::PreparedSurfaceSIF, ::Lambertian...) (Sources/surface_sif.jl:168-178)
            j0m_s = j0m_s + FT(2) * np.tile(np.asarray(model.sif0, FT)[idx], (1, q.Nquad))
        C = interaction(ctl.interface[-1], C, r_mp_s, r_pm_s, t_s, t_mm_s, j0p_s, j0m_s)
        # interaction_hdrf!, interaction_hdrf.jl:4-42
        hdr_J0m = _mv(r_mp_s, C.J0p) + j0m_s
        if m == 0:
            wq = (q.wt_muN * q.qp_muN).astype(FT)
            i0 = pol_n * q.imu0
            for i in range(pol_n):
                j = slice(i, None, pol_n)
                bhr_uw[i, :] = np.sum(hdr_J0m[:, j] * wq[j][None, :], axis=1)
                bhr_dw[i, :] = (np.sum(C.J0p[:, j] * wq[j][None, :], axis=1)
                                + j0p_s[:, i0] * q.qp_muN[i0])
        # postprocessing_vza!, postprocessing_vza.jl:23-94
        for i in range(nV):
            imu = nearest_point(q.qp_mu, FT(cosd(model.vza[i])))
            ist = imu * pol_n
            c = cosd(m * model.vaz[i])
            sn = sind(m * model.vaz[i])
            wv = (weight * np.array([c, c, sn, sn][:pol_n])).astype(FT)
            R_SFI[i] += wv[:, None] * C.J0m[:, ist:ist + pol_n].T
            T_SFI[i] += wv[:, None] * C.J0p[:, ist:ist + pol_n].T
            hdr[i] += wv[:, None] * hdr_J0m[:, ist:ist + pol_n].T
        if return_streams:
            streams.append(C)
    if model.ss_correction is not None:                 # apply_ss_correction! (coxmunk_surface.jl:544-552)
        ctl0 = prep.ctl if prep is not None else band_control(model, 0)
        atten = (FT(q.mu0) * np.exp(-ctl0.tau_sum[idx, -1] / FT(q.mu0))).astype(FT)
        R_SFI += np.asarray(model.ss_correction, FT)[:, :, None] * atten[None, None, :]
    out = dict(R_SFI=R_SFI, T_SFI=T_SFI, hdr=hdr, bhr_uw=bhr_uw[0], bhr_dw=bhr_dw[0])
    if return_streams:
        out['streams'] = streams
    return out


# --------------------------------------------------------------------------
# Linearized column (config 4 subset): rt_run(model, lin_model, ...) of rt_run_lin.jl:102-326
# for parameters that act through the layer optical depths and the Lambertian albedo.
# --------------------------------------------------------------------------
def core_props_lin(model: Model, m: int, dtau_rayl, dtau_abs, dtau_aer, ztab=None):
    """Per-layer derivatives (tau_dot, varpi_dot, Zpp_dot, Zmp_dot) of the mixed layer properties
    (constructCoreOpticalProperties of LinMode, compEffectiveLayerProperties_lin.jl) for P
    parameters given as derivatives of the inputs: dtau_rayl, dtau_abs (P,nSpec,Nz), dtau_aer
    (P,nAer,Nz).  With S = sum_c tau_c varpi_c:  varpi = S/tau,  Z = sum_c tau_c varpi_c Z_c / S."""
    FT = model.FT
    (Zpp_r, Zmp_r), aerZ = ztab if ztab is not None else z_tables(model, m)
    P = dtau_abs.shape[0]
    out = []
    for iz in range(model.Nz):
        comps = [(model.tau_rayl[:, iz].astype(FT), FT(model.varpi_cabannes), Zpp_r, Zmp_r, dtau_rayl[:, :, iz])]
        for ia, a in enumerate(model.aerosols):
            k = FT(1 - a.ft * a.omega)
            comps.append((np.full(model.nSpec, k * model.tau_aer[ia, iz], FT), FT((1 - a.ft) * a.omega / (1 - a.ft * a.omega)),
                          aerZ[ia][0], aerZ[ia][1], np.broadcast_to((k * dtau_aer[:, ia, iz])[:, None], (P, model.nSpec))))
        tau = sum(c[0] for c in comps) + model.tau_abs[:, iz].astype(FT)
        S = sum(c[0] * c[1] for c in comps)
        dtau = sum(c[4] for c in comps) + dtau_abs[:, :, iz]
        dS = sum(c[4] * c[1] for c in comps)
        safe_tau = np.where(tau > 0, tau, 1)
        safe_S = np.where(S > 0, S, 1)
        varpi = S / safe_tau
        dvarpi = (dS - varpi[None] * dtau) / safe_tau[None]
        Zpp = sum((c[0] * c[1])[:, None, None] * c[2][None] for c in comps) / safe_S[:, None, None]
        Zmp = sum((c[0] * c[1])[:, None, None] * c[3][None] for c in comps) / safe_S[:, None, None]
        dZpp = (sum((c[4] * c[1])[:, :, None, None] * c[2][None, None] for c in comps) - dS[:, :, None, None] * Zpp[None]) / safe_S[None, :, None, None]
        dZmp = (sum((c[4] * c[1])[:, :, None, None] * c[3][None, None] for c in comps) - dS[:, :, None, None] * Zmp[None]) / safe_S[None, :, None, None]
        out.append((dtau, dvarpi, dZpp, dZmp))
    return out


def rt_run_lin(model: Model, dtau_rayl, dtau_abs, dtau_aer, dalbedo, ctl: Optional[BandControl] = None):
    """Forward result and Jacobian dR_SFI/dp, dT_SFI/dp (P,nVZA,pol_n,nSpec) for P parameters that
    perturb tau_rayl, tau_abs (P,nSpec,Nz), tau_aer (P,nAer,Nz) and the Lambertian albedo (P,).
    The layer loop of rt_run_lin.jl:102-326 with rt_kernel! of rt_kernel_lin.jl:50-180; control
    flow (ndoubl, interfaces) is that of the unperturbed band, as in the reference; `ctl` overrides it (a sample of a
    larger band must branch like the full band)."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    nS, nV, N = model.nSpec, len(model.vza), q.qp_muN.size
    P = dtau_abs.shape[0]
    R = np.zeros((nV, pol_n, nS), FT); T = np.zeros_like(R)
    dR = np.zeros((P, nV, pol_n, nS), FT); dT = np.zeros_like(dR)
    F0 = model.F0
    if F0 is None:
        F0 = np.zeros((pol_n, nS), FT); F0[0, :] = 1
    ctl_given = ctl
    for m in range(model.m_max + 1):
        weight = FT(0.5 / math.pi) if m == 0 else FT(1.0 / math.pi)
        ztab = z_tables(model, m)
        if ctl_given is None:
            ctl = band_control(model, m, ztab)
        props = construct_core_optical_properties(model, m, ztab=ztab)
        lprops = core_props_lin(model, m, dtau_rayl, dtau_abs, dtau_aer, ztab)
        dtausum = np.zeros((P, nS), FT)
        C = dC = None
        for iz in range(model.Nz):
            pr = props[iz]
            dtl, dvl, dZp, dZm = lprops[iz]
            if ctl.scatter[iz]:
                nd = ctl.ndoubl[iz]
                sc = FT(2 ** nd)
                dtau = (pr.tau / sc).astype(FT)
                expk = np.exp(-dtau / FT(q.mu0))
                Zpp = np.broadcast_to(pr.Zpp, (nS, N, N)); Zmp = np.broadcast_to(pr.Zmp, (nS, N, N))
                a = elemental(model, m, nd, dtau, pr.varpi, Zpp, Zmp, ctl.tau_sum[:, iz], F0)
                da = elemental_lin(model, m, nd, dtau, pr.varpi, Zpp, Zmp, ctl.tau_sum[:, iz], F0,
                                   dtl / sc, dvl, dZp, dZm, dtausum)
                if nd > 0:
                    a, da = doubling_lin(model, nd, expk, a[0], a[1], a[4], a[5], da[0], da[1], da[4], da[5], dtl / sc)
                r, t, r_pm, t_mm, j0p, j0m = a
                dr, dt, dr_pm, dt_mm, dj0p, dj0m = da
            else:       # zero_added_noscat! and its derivative: t = diag(exp(-tau/mu))
                tb = np.exp(-pr.tau[:, None] / q.qp_muN.astype(FT)[None, :])
                idx = np.arange(N)
                r = np.zeros((nS, N, N), FT); r_pm = r.copy(); t = r.copy(); t[:, idx, idx] = tb; t_mm = t.copy()
                j0p = np.zeros((nS, N), FT); j0m = j0p.copy()
                dr = np.zeros((P, nS, N, N), FT); dr_pm = dr.copy(); dt = dr.copy()
                dt[:, :, idx, idx] = -(tb / q.qp_muN.astype(FT)[None, :])[None] * dtl[:, :, None]
                dt_mm = dt.copy()
                dj0p = np.zeros((P, nS, N), FT); dj0m = dj0p.copy()
            if iz == 0:
                C = Composite(r.copy(), r_pm.copy(), t.copy(), t_mm.copy(), j0p.copy(), j0m.copy())
                dC = Composite(dr.copy(), dr_pm.copy(), dt.copy(), dt_mm.copy(), dj0p.copy(), dj0m.copy())
            else:
                C, dC = interaction_lin(ctl.interface[iz], C, dC, r, r_pm, t, t_mm, j0p, j0m, dr, dr_pm, dt, dt_mm, dj0p, dj0m)
            dtausum = dtausum + dtl
        # surface and its derivative: Lambertian (lambertian_surface_lin.jl: albedo and attenuation), or a BRDF given by
        # its Fourier moments (no surface parameter here: only the attenuation of the direct beam varies)
        brdf = model.surface_rho is not None
        if brdf:
            r_s, rpm_s, t_s, tmm_s, j0p_s, j0m_s = brdf_surface(model, m, ctl.tau_sum[:, -1])
        else:
            r_s, rpm_s, t_s, tmm_s, j0p_s, j0m_s = lambertian_surface(model, m, ctl.tau_sum[:, -1])
        dr_s = np.zeros((P, nS, N, N), FT); dj0p_s = np.zeros((P, nS, N), FT); dj0m_s = np.zeros((P, nS, N), FT)
        if m == 0 or brdf:
            att_d = -(1 / FT(q.mu0)) * dtausum                                  # d att / att
            dj0p_s = j0p_s[None] * att_d[:, :, None]
            dj0m_s = j0m_s[None] * att_d[:, :, None]
            if not brdf and model.albedo != 0:
                rel = (np.asarray(dalbedo, FT) / FT(model.albedo))[:, None, None]
                dr_s = dr_s + r_s[None] * rel[:, :, :, None]
                dj0m_s = dj0m_s + j0m_s[None] * rel
        z4 = np.zeros((P, nS, N, N), FT)
        C, dC = interaction_lin(ctl.interface[-1], C, dC, r_s, rpm_s, t_s, tmm_s, j0p_s, j0m_s,
                                dr_s, z4, z4.copy(), z4.copy(), dj0p_s, dj0m_s)
        for i in range(nV):      # postprocessing_vza! (lin: postprocessing_vza_lin.jl:1-48)
            imu = nearest_point(q.qp_mu, FT(cosd(model.vza[i])))
            ist = imu * pol_n
            c, sn = cosd(m * model.vaz[i]), sind(m * model.vaz[i])
            wv = (weight * np.array([c, c, sn, sn][:pol_n])).astype(FT)
            R[i] += wv[:, None] * C.J0m[:, ist:ist + pol_n].T
            T[i] += wv[:, None] * C.J0p[:, ist:ist + pol_n].T
            dR[:, i] += wv[None, :, None] * np.swapaxes(dC.J0m[:, :, ist:ist + pol_n], 1, 2)
            dT[:, i] += wv[None, :, None] * np.swapaxes(dC.J0p[:, :, ist:ist + pol_n], 1, 2)
    return dict(R_SFI=R, T_SFI=T, dR_SFI=dR, dT_SFI=dT)


def rt_run_threaded(model: Model, idx: np.ndarray, threads: int, chunk: int = 8,
                    prep: Optional[Prepared] = None):
    """CPU baseline driver: the reference threads its per-slice BLAS/LAPACK calls
    over the spectral axis (cpu_batched.jl:26,33,78); here the selected spectral
    points are split into chunks processed by a thread pool (NumPy releases the
    GIL inside matmul / LAPACK; BLAS itself is pinned to one thread by the caller).
    Returns R_SFI, T_SFI for the selected points."""
    from concurrent.futures import ThreadPoolExecutor
    if prep is None:
        prep = prepare(model)
    idx = np.asarray(idx)
    parts = [idx[i:i + chunk] for i in range(0, idx.size, chunk)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        outs = list(ex.map(lambda sl: rt_run(model, sl, prep=prep), parts))
    return (np.concatenate([o["R_SFI"] for o in outs], axis=2),
            np.concatenate([o["T_SFI"] for o in outs], axis=2))


_MP_STATE = {}


def _mp_worker(part):
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            o = rt_run(_MP_STATE["model"], part, prep=_MP_STATE["prep"])
    except ImportError:
        o = rt_run(_MP_STATE["model"], part, prep=_MP_STATE["prep"])
    return o["R_SFI"], o["T_SFI"]


class ProcessPoolBaseline:
    """CPU baseline over all host cores: one forked worker per core, each running
    the oracle on chunks of the spectral axis with BLAS pinned to one thread -- the
    reference's own CPU layout (Threads.@threads over nSpec with single-threaded
    per-slice gemm / getrf, cpu_batched.jl:26-37,65-82)."""

    def __init__(self, model: Model, workers: int, prep: Optional[Prepared] = None):
        import multiprocessing as mp
        _MP_STATE["model"] = model
        _MP_STATE["prep"] = prep if prep is not None else prepare(model)
        self.workers = workers
        self.pool = mp.get_context("fork").Pool(workers)

    def run(self, idx: np.ndarray, chunk: int = 8):
        idx = np.asarray(idx)
        parts = [idx[i:i + chunk] for i in range(0, idx.size, chunk)]
        outs = self.pool.map(_mp_worker, parts, chunksize=1)
        return (np.concatenate([o[0] for o in outs], axis=2), np.concatenate([o[1] for o in outs], axis=2))

    def close(self):
        self.pool.close()
        self.pool.join()


# --------------------------------------------------------------------------
# Single-scattering driver: rt_run_ss(::noRS, model, iBand) rt_run.jl:593-755 with rt_kernel_ss!(::noRS{FT})
# CoreKernel/rt_kernel_ss.jl:156-236 (elemental! of the WHOLE layer with ndoubl = 0, no doubling) and
# interaction_ss! CoreKernel/interaction_ss.jl:14-42.  Pinned by the invariants of test/test_forward_ss.jl (finite,
# positive, SS <= MS on Stokes I, hemispheric bounds) and by the thin-atmosphere limit R_ss -> R_ms
# (tests/test_oracle_golden.py::test_single_scatter_*).
# --------------------------------------------------------------------------
def rt_run_ss(model: Model):
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    nS = model.nSpec
    nV = len(model.vza)
    mu = q.qp_muN.astype(FT)
    R_SFI = np.zeros((nV, pol_n, nS), FT)
    T_SFI = np.zeros((nV, pol_n, nS), FT)
    hem_R = np.zeros(nS, FT)
    hem_T = np.zeros(nS, FT)
    F0 = model.F0
    if F0 is None:
        F0 = np.zeros((pol_n, nS), FT)
        F0[0, :] = 1
    for m in range(model.m_max + 1):
        weight = FT(0.5 / math.pi) if m == 0 else FT(1.0 / math.pi)
        ctl = band_control(model, m)
        props = construct_core_optical_properties(model, m)
        J0p = J0m = None
        for iz in range(model.Nz):
            p = props[iz]
            tau = np.broadcast_to(p.tau, (nS,)).astype(FT)
            varpi = np.broadcast_to(p.varpi, (nS,)).astype(FT)
            tsum = ctl.tau_sum[:, iz]
            _, _, _, _, j0p, j0m = elemental(model, m, 0, tau, varpi, p.Zpp, p.Zmp, tsum, F0)   # ndoubl_ss = 0
            if iz == 0:
                J0p, J0m = j0p.copy(), j0m.copy()
            else:                                                  # get_interaction_ss! (interaction_ss.jl:14-19)
                J0p = J0p * np.exp(-tau[:, None] / mu[None, :]) + j0p
                J0m = J0m + j0m * np.exp(-tsum[:, None] / mu[None, :])
        s = brdf_surface(model, m, ctl.tau_sum[:, -1]) if model.surface_rho is not None \
            else lambertian_surface(model, m, ctl.tau_sum[:, -1])
        j0p_s, j0m_s = s[4], s[5]
        J0p = J0p + j0p_s                                          # τ_surf = 0 (rt_run.jl:713-721)
        J0m = J0m + j0m_s * np.exp(-ctl.tau_sum[:, -1][:, None] / mu[None, :])
        for i in range(nV):
            imu = nearest_point(q.qp_mu, FT(cosd(model.vza[i])))
            ist = imu * pol_n
            c = cosd(m * model.vaz[i])
            sn = sind(m * model.vaz[i])
            wv = (weight * np.array([c, c, sn, sn][:pol_n])).astype(FT)
            R_SFI[i] += wv[:, None] * J0m[:, ist:ist + pol_n].T
            T_SFI[i] += wv[:, None] * J0p[:, ist:ist + pol_n].T
        if m == 0:                                                 # rt_run.jl:737-748
            wq = (q.wt_mu * q.qp_mu).astype(FT)
            hem_R += np.sum(J0m[:, 0::pol_n] * wq[None, :], axis=1)
            hem_T += np.sum(J0p[:, 0::pol_n] * wq[None, :], axis=1)
    return dict(R_SFI=R_SFI, T_SFI=T_SFI, hem_R=hem_R, hem_T=hem_T)

"""Cox-Munk ocean surface for the B200 path -- host-side mirror of
``src/CoreRT/Surfaces/coxmunk_surface.jl`` (+ ``fresnel.jl``): the reference builds the Fourier
moments of the polarized glint BRDF on the host, once per band (``reflectance(::CoxMunkSurface, ...)``,
coxmunk_surface.jl:464-498), and applies the TMS single-scattering correction on the host after the
Fourier loop (``apply_ss_correction!``, :509-555).  Here both are NumPy-vectorised over
(reflected stream, incident stream, azimuth node); the device consumes the moments through
``RTModel.surface_rho`` and adds the correction in its output reduction (``RTModel.ss_correction``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy.special import erfc

# water_refractive_index(550 nm): the 550 nm row of the Segelstein (1981) table (water_refraction.jl).  All
# reference call sites use the default `n_water = _get_n_water(surf, FT(550))` (coxmunk_surface.jl:262,465,514).
N_WATER_550 = complex(1.328, 5.69e-9)


@dataclass
class CoxMunkSurface:
    """``CoxMunkSurface`` (src/CoreRT/types.jl:525-536)."""
    wind_speed: float
    n_water: complex = N_WATER_550
    whitecap_albedo: float = 0.22
    include_whitecaps: bool = True
    shadowing: bool = True

    @property
    def sigma2(self) -> float:                      # wind_to_sigma2, coxmunk_surface.jl:22
        return 0.003 + 0.00512 * self.wind_speed

    @property
    def whitecap_fraction(self) -> float:           # :30-33
        return 0.0 if self.wind_speed <= 0 else 2.95e-6 * self.wind_speed ** 3.52


def _smith_lambda(mu, s2):
    """Smith (1967) shadowing integral (coxmunk_surface.jl:41-50), array version."""
    mu = np.asarray(mu, float)
    cot = mu / np.sqrt(np.maximum(1e-30, 1.0 - mu * mu))
    nu = np.where(mu > 0, cot, 1.0) / math.sqrt(2.0 * s2)
    lam = np.maximum(0.0, (np.exp(-nu * nu) / (math.sqrt(2 * math.pi) * nu) - erfc(nu)) / 2)
    return np.where(mu > 0, lam, 1e10)


def _signed_angle(dot, m1, m2, sign):
    ok = ~((m1 < 1e-15) | (m2 < 1e-15))
    c = np.clip(dot / np.where(ok, m1 * m2, 1.0), -1.0, 1.0)
    a = np.arccos(c)
    return np.where(ok, np.where(sign >= 0, a, -a), 0.0)


def brdf_mueller(surf: CoxMunkSurface, n: int, mu_r, mu_i, dphi) -> np.ndarray:
    """``coxmunk_brdf_mueller(surf, n, mu[i], mu[j], dphi)`` (coxmunk_surface.jl:259-301) for broadcastable
    arrays: first direction = the row (reflected) stream, second = the column (incident) stream, exactly the
    argument order of the reference's callers (:482, :520).  Returns (..., n, n)."""
    a, b, dphi = np.broadcast_arrays(np.asarray(mu_r, float), np.asarray(mu_i, float), np.asarray(dphi, float))
    # coxmunk_geometry (:124-250) names its arguments (mu_i, mu_r) = (a, b)
    sa = np.sqrt(np.maximum(0.0, 1 - a * a))
    sb = np.sqrt(np.maximum(0.0, 1 - b * b))
    cd, sd = np.cos(dphi), np.sin(dphi)
    nx = -sa + sb * cd
    ny = sb * sd
    nz = a + b
    nn = np.sqrt(nx * nx + ny * ny + nz * nz)
    deg = nn < 1e-15
    nns = np.where(deg, 1.0, nn)
    nx, ny, nz = nx / nns, ny / nns, nz / nns
    cos_b = np.maximum(1e-10, nz)
    cos_loc = np.clip((a + b) / (2 * cos_b), 0.0, 1.0)
    zx, zy = -nx / cos_b, -ny / cos_b
    cosT = -a * b + sa * sb * cd
    sinT = np.sqrt(np.maximum(0.0, 1 - cosT * cosT))
    # plane normals: scattering plane (k_i x k_r), incidence plane (k_i x n), reflection plane (k_r x n)
    sp = np.stack([-a * sb * sd, a * sb * cd - sa * b, -sa * sb * sd], -1)
    ip = np.stack([a * ny, -a * nx - sa * nz, sa * ny], -1)
    rp = np.stack([(-sb * sd) * nz - b * ny, b * nx + sb * cd * nz, (-sb * cd) * ny + sb * sd * nx], -1)
    msp = np.linalg.norm(sp, axis=-1)
    c1 = np.cross(sp, ip)
    a1 = _signed_angle((sp * ip).sum(-1), msp, np.linalg.norm(ip, axis=-1), sa * c1[..., 0] - a * c1[..., 2])
    c2 = np.cross(sp, rp)
    a2 = _signed_angle((sp * rp).sum(-1), msp, np.linalg.norm(rp, axis=-1),
                       (-sb * cd) * c2[..., 0] + (-sb * sd) * c2[..., 1] + b * c2[..., 2])
    flat = sinT < 1e-12
    a1 = np.where(flat | deg, 0.0, a1)
    a2 = np.where(flat | deg, 0.0, a2)
    cos_b = np.where(deg, 1.0, cos_b); cos_loc = np.where(deg, 1.0, cos_loc)
    zx = np.where(deg, 0.0, zx); zy = np.where(deg, 0.0, zy)
    s2 = surf.sigma2
    P = np.exp(-(zx * zx + zy * zy) / (2 * s2)) / (2 * math.pi * s2)
    # Fresnel amplitudes (fresnel.jl:25-33) and Mueller matrix (:50-88)
    nw = complex(surf.n_water)
    cos_t = np.sqrt(1.0 - np.maximum(0.0, 1 - cos_loc * cos_loc) / (nw * nw) + 0j)
    r_s = (cos_loc - nw * cos_t) / (cos_loc + nw * cos_t)
    r_p = (nw * cos_loc - cos_t) / (nw * cos_loc + cos_t)
    rs2, rp2 = np.abs(r_s) ** 2, np.abs(r_p) ** 2
    rsp = r_s * np.conj(r_p)
    F = np.zeros(a.shape + (4, 4))
    F[..., 0, 0] = F[..., 1, 1] = (rs2 + rp2) / 2
    F[..., 0, 1] = F[..., 1, 0] = (rs2 - rp2) / 2
    F[..., 2, 2] = F[..., 3, 3] = rsp.real
    F[..., 2, 3] = rsp.imag
    F[..., 3, 2] = -rsp.imag

    def rot(phi):                                   # stokes_rotation_matrix (fresnel.jl:104-132)
        L = np.zeros(a.shape + (4, 4))
        L[..., 0, 0] = L[..., 3, 3] = 1.0
        L[..., 1, 1] = L[..., 2, 2] = np.cos(2 * phi)
        L[..., 1, 2] = np.sin(2 * phi)
        L[..., 2, 1] = -np.sin(2 * phi)
        return L[..., :n, :n]
    M = rot(a2) @ F[..., :n, :n] @ rot(-a1)
    pre = P / (4 * a * b * cos_b ** 4)
    if surf.shadowing:
        pre = pre / (1.0 + _smith_lambda(a, s2) + _smith_lambda(b, s2))
    M = pre[..., None, None] * M
    if surf.include_whitecaps:
        f = surf.whitecap_fraction
        M = (1 - f) * M
        M[..., 0, 0] += f * surf.whitecap_albedo / math.pi
    return M


def _gauleg_0_pi(nquad: int):
    x, w = np.polynomial.legendre.leggauss(nquad)
    return 0.5 * math.pi * (x + 1.0), 0.5 * math.pi * w


def _az_cos_mask(n: int) -> np.ndarray:
    """True where the azimuthal kernel of Stokes element (si, sj) is cos(m dphi), False where sin
    (_azimuthal_kernel, coxmunk_surface.jl:444-452)."""
    iq = np.arange(n) <= 1
    return iq[:, None] == iq[None, :]


def coxmunk_surface_rho(quad, pol_n: int, m_max: int, surf: CoxMunkSurface, nquad: int = 100) -> np.ndarray:
    """Fourier moments 0..m_max of the Cox-Munk reflectance matrix as ``create_surface_layer!`` uses them
    (rpv_surface.jl:69-73 on top of ``reflectance``, coxmunk_surface.jl:464-498): full (pol_n x pol_n) Mueller
    blocks, 100-point Gauss-Legendre rule over [0, pi], overall factor 2/pi.  Shape (m_max+1, N, N), ready for
    ``RTModel.surface_rho``."""
    mu = np.asarray(quad.qp_mu, np.float64)
    nq = mu.size
    phi, w = _gauleg_0_pi(nquad)
    M = brdf_mueller(surf, pol_n, mu[:, None, None], mu[None, :, None], phi[None, None, :])   # (i, j, phi, si, sj)
    cosm = _az_cos_mask(pol_n)
    out = np.empty((m_max + 1, nq * pol_n, nq * pol_n))
    for m in range(m_max + 1):
        c = np.einsum("ijpab,p->ijab", M, w * np.cos(m * phi))
        s = np.einsum("ijpab,p->ijab", M, w * np.sin(m * phi))
        blk = np.where(cosm[None, None], c, s)                      # (i, j, si, sj)
        out[m] = (2.0 / math.pi) * blk.transpose(0, 2, 1, 3).reshape(nq * pol_n, nq * pol_n)
    return out


def coxmunk_ss_correction(surf: CoxMunkSurface, pol_n: int, vza, vaz, mu0: float, m_max: int,
                          nquad: int = 100) -> np.ndarray:
    """Geometry factor of ``apply_ss_correction!`` (coxmunk_surface.jl:509-555): (nVZA, pol_n) array
    ``M_exact[si,1] - M_fourier[si,1]`` with the truncated reconstruction summed over m = 0..m_max
    (weights 0.5 / 1, the same quadrature as the moments).  The TMS correction is then
    ``R_SFI[iv,si,s] += mu0 * exp(-tau_total[s]/mu0) * factor[iv,si]``."""
    mu_v = np.cos(np.radians(np.asarray(vza, float)))
    dphi = np.radians(np.asarray(vaz, float))
    phi, w = _gauleg_0_pi(nquad)
    exact = brdf_mueller(surf, pol_n, mu_v, mu0, dphi)[..., :, 0]                         # (nV, si)
    col = brdf_mueller(surf, pol_n, mu_v[:, None], mu0, phi[None, :])[..., :, 0]          # (nV, phi, si)
    cos_row = np.arange(pol_n) <= 1                     # kernel of (si, sj = I)
    four = np.zeros_like(exact)
    for m in range(m_max + 1):
        ff = 1.0 if m == 0 else 2.0
        wm = 0.5 if m == 0 else 1.0
        kq = np.where(cos_row[None, :], np.cos(m * phi)[:, None], np.sin(m * phi)[:, None])      # (phi, si)
        coeff = ff * np.einsum("vps,p,ps->vs", col, w, kq) / math.pi
        kv = np.where(cos_row[None, :], np.cos(m * dphi)[:, None], np.sin(m * dphi)[:, None])    # (nV, si)
        four += wm * kv * coeff
    return exact - four

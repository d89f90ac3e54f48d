"""Cox-Munk ocean surface: the oracle restatement pinned to the reference's own known-answer tests
(test/test_coxmunk.jl -- the reference holds no numeric golden table for this surface, only these
analytic checks and invariants), and the vectorised host mirror (vsmom_b200.coxmunk) against the oracle."""
import math

import numpy as np

import vsmom_b200 as vb
import vsmom_oracle as orc
from vsmom_b200 import coxmunk as cm

NW = complex(1.33, 0.0)


def test_fresnel_coefficients_kat():
    # test_coxmunk.jl:29-59
    rs, rp = orc.fresnel_coefficients(NW, 1.0)
    assert abs(abs(rs) ** 2 - ((1.33 - 1) / (1.33 + 1)) ** 2) < 1e-10
    assert abs(abs(rs) ** 2 - abs(rp) ** 2) < 1e-12
    rs, rp = orc.fresnel_coefficients(NW, math.cos(math.atan(1.33)))          # Brewster angle
    assert abs(rp) < 1e-10
    rs, rp = orc.fresnel_coefficients(NW, 0.001)                               # grazing
    assert abs(rs) ** 2 > 0.99 and abs(rp) ** 2 > 0.99
    rs, rp = orc.fresnel_coefficients(complex(1.33, 0.01), 0.5)
    assert math.isfinite(abs(rs) ** 2) and math.isfinite(abs(rp) ** 2)


def test_fresnel_mueller_and_rotation_kat():
    # test_coxmunk.jl:64-137
    rs, rp = orc.fresnel_coefficients(NW, 0.7)
    M = orc.fresnel_mueller(rs, rp, 4)
    assert np.all(np.abs(M[:2, 2:]) < 1e-15) and np.all(np.abs(M[2:, :2]) < 1e-15)
    assert M[0, 1] == M[1, 0] and M[2, 2] == M[3, 3] and M[2, 3] == -M[3, 2]
    assert M[0, 0] >= abs(M[0, 1]) and M[0, 0] <= 1.0
    for n in (1, 2, 3, 4):
        Mn = orc.fresnel_mueller(rs, rp, n)
        assert Mn.shape == (n, n) and np.allclose(Mn, M[:n, :n])
    assert np.allclose(orc.stokes_rotation_matrix(0.0, 4), np.eye(4), atol=1e-15)
    a, b = 0.3, 0.5
    La, Lb = orc.stokes_rotation_matrix(a, 4), orc.stokes_rotation_matrix(b, 4)
    assert np.allclose(La @ Lb, orc.stokes_rotation_matrix(a + b, 4), atol=1e-12)
    assert np.allclose(La @ orc.stokes_rotation_matrix(-a, 4), np.eye(4), atol=1e-12)
    assert np.allclose(La.T @ La, np.eye(4), atol=1e-12)
    # docstring layout (fresnel.jl:95-101): L[2,3] = sin 2phi, L[3,2] = -sin 2phi (1-based)
    assert abs(La[1, 2] - math.sin(2 * a)) < 1e-15 and abs(La[2, 1] + math.sin(2 * a)) < 1e-15


def test_helpers_kat():
    # test_coxmunk.jl:173-215
    assert abs(orc.wind_to_sigma2(0.0) - 0.003) < 1e-10
    assert abs(orc.wind_to_sigma2(10.0) - (0.003 + 0.0512)) < 1e-10
    assert orc.whitecap_fraction(0.0) == 0.0
    assert abs(orc.whitecap_fraction(10.0) / (2.95e-6 * 10.0 ** 3.52) - 1) < 1e-6
    s2 = orc.wind_to_sigma2(5.0)
    z = np.linspace(-1.5, 1.5, 301)
    dz = z[1] - z[0]
    integral = sum(orc.cox_munk_pdf(zx, zy, s2) for zx in z for zy in z) * dz * dz
    assert abs(integral - 1.0) < 0.01
    assert orc.shadow_factor(0.95, 0.95, s2) > 0.9
    assert 0.0 < orc.shadow_factor(0.1, 0.1, s2) < orc.shadow_factor(0.95, 0.95, s2)


def test_brdf_mueller_kat():
    # test_coxmunk.jl:219-287, 353-375
    surf = orc.CoxMunk(5.0, NW, include_whitecaps=False)
    M = orc.coxmunk_brdf_mueller(surf, 4, 0.6, 0.8, 1.2)
    assert np.isfinite(M).all()
    Mr = orc.coxmunk_brdf_mueller(surf, 4, 0.8, 0.6, 1.2)
    assert abs(M[0, 0] / Mr[0, 0] - 1) < 1e-6                                   # reciprocity
    lo = orc.coxmunk_brdf_mueller(orc.CoxMunk(1.0, NW, include_whitecaps=False), 1, 0.7, 0.7, 0.0)
    hi = orc.coxmunk_brdf_mueller(orc.CoxMunk(20.0, NW, include_whitecaps=False), 1, 0.7, 0.7, 0.0)
    assert lo[0, 0] > 5 * hi[0, 0]                                              # specular peak sharpens at low wind
    n_mu, n_phi = 30, 60
    tot = 0.0
    for i in range(n_mu):
        mu_out = (i + 0.5) / n_mu
        for j in range(n_phi):
            tot += orc.coxmunk_brdf_mueller(surf, 1, mu_out, 0.5, (j + 0.5) * 2 * math.pi / n_phi)[0, 0] * mu_out
    assert tot / n_mu * 2 * math.pi / n_phi <= 1.02                             # energy conservation
    wc = orc.coxmunk_brdf_mueller(orc.CoxMunk(12.0, NW), 1, 0.3, 0.8, 2.0)
    no = orc.coxmunk_brdf_mueller(orc.CoxMunk(12.0, NW, include_whitecaps=False), 1, 0.3, 0.8, 2.0)
    assert wc[0, 0] != no[0, 0]
    M = orc.coxmunk_brdf_mueller(orc.CoxMunk(3.0, NW, include_whitecaps=False), 4, 0.7, 0.7, 0.0)
    assert M[0, 0] > 0 and abs(M[0, 1] - M[1, 0]) < 1e-10 and abs(M[0, 1]) / M[0, 0] > 0.01


def _oq(q, n):
    return orc.QuadPoints(q.mu0, q.imu0 - 1, q.qp_mu, q.wt_mu, q.qp_muN, q.wt_muN, q.Nquad, q.Nstreams, n)


def test_fourier_decomposition_kat():
    # test_coxmunk.jl:290-349: sizes, finiteness, positivity of the scalar m = 0 moment, decay with m, I-Q coupling
    surf = orc.CoxMunk(5.0, NW)
    mu = np.array([0.3, 0.5, 0.7, 0.9])
    for n in (1, 2, 3, 4):
        q = orc.QuadPoints(0.5, 1, mu, np.ones(4) / 4, np.repeat(mu, n), np.repeat(np.ones(4) / 4, n), 4, 4, n)
        R = orc.coxmunk_surface_rho(q, 10 if n == 1 else 0, surf, nquad=100)
        assert R.shape[1:] == (4 * n, 4 * n) and np.isfinite(R).all()
        if n == 1:
            assert (R[0] >= -1e-10).all()
            assert np.abs(R[10]).max() / 2 < np.abs(R[0]).max() / 2 + 1e-300 or np.abs(R[10]).max() / 2 < np.abs(R[0]).max()
        if n == 4:
            assert np.abs(R[0][1::4, 0::4]).max() > 1e-12


def test_mirror_matches_oracle_pointwise_and_moments():
    so, sm = orc.CoxMunk(5.0), cm.CoxMunkSurface(5.0)
    assert so.n_water == sm.n_water == complex(1.328, 5.69e-9)
    rng = np.random.default_rng(3)
    for _ in range(200):
        a, b = rng.uniform(0.02, 1.0, 2)
        d = rng.uniform(0.0, math.pi)
        A = orc.coxmunk_brdf_mueller(so, 4, a, b, d)
        B = cm.brdf_mueller(sm, 4, a, b, d)
        assert np.abs(A - B).max() <= 1e-12 * max(np.abs(A).max(), 1e-300)
    for g in ((0.7, 0.7, 0.0), (1.0, 1.0, 0.3), (0.5, 0.5, math.pi), (1.0, 0.5, 0.0)):     # degenerate geometries
        assert np.abs(orc.coxmunk_brdf_mueller(so, 4, *g) - cm.brdf_mueller(sm, 4, *g)).max() < 1e-15
    q = vb.rt_set_streams(3, 30.0, [60.0, 0.0], 3)
    R1 = orc.coxmunk_surface_rho(_oq(q, 3), 5, so)
    R2 = cm.coxmunk_surface_rho(q, 3, 5, sm)
    assert np.abs(R1 - R2).max() <= 1e-9 * np.abs(R1).max()
    c1 = orc.coxmunk_ss_correction(so, 3, [60.0, 0.0, 30.0], [180.0, 0.0, 0.0], q.mu0, 5)
    c2 = cm.coxmunk_ss_correction(sm, 3, [60.0, 0.0, 30.0], [180.0, 0.0, 0.0], q.mu0, 5)
    assert np.abs(c1 - c2).max() <= 1e-12


def test_ocean_scene_shapes_and_oracle_correction():
    # BASELINE.json configs[2] shape: 20 streams + 5 zero-weight nodes -> N = 75 (IQU) / 100 (IQUV), 40 moments
    m = vb.synthetic_ocean_model(nSpec=4, nstreams=20, polarization_type="Stokes_IQUV")
    assert m.quad_points.qp_muN.size == 100 and m.m_max == 39 and m.Nz == 33
    assert m.surface_rho.shape == (40, 100, 100) and m.ss_correction.shape == (9, 4)
    m = vb.synthetic_ocean_model(nSpec=2, nstreams=3, polarization_type="Stokes_IQU")
    from helpers import to_oracle
    on = orc.rt_run(to_oracle(m))
    m.ss_correction = None
    off = orc.rt_run(to_oracle(m))
    tau = (m.tau_rayl + m.tau_abs).sum(axis=1) + m.tau_aer.sum()
    mu0 = m.quad_points.mu0
    m2 = vb.synthetic_ocean_model(nSpec=2, nstreams=3, polarization_type="Stokes_IQU")
    want = m2.ss_correction[:, :, None] * (mu0 * np.exp(-tau / mu0))[None, None, :]
    assert np.allclose(on["R_SFI"] - off["R_SFI"], want, rtol=1e-9, atol=1e-16)
    assert np.array_equal(on["T_SFI"], off["T_SFI"])

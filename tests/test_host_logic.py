"""CPU tests of the host-side mirror (vsmom_b200.model) against the oracle's
independent restatement: quadrature, Z moments, full-band control flow."""
import numpy as np
import pytest

import vsmom_b200 as vb
import vsmom_oracle as orc
from helpers import to_oracle


@pytest.mark.parametrize("pol", ["Stokes_I", "Stokes_IQU", "Stokes_IQUV"])
def test_streams_match_oracle(pol):
    n = vb.model.POL_N[pol]
    q = vb.rt_set_streams(10, 60.0, vb.model.DEFAULT_VZA, n)
    o = orc.rt_set_streams_gauss(10, 60.0, vb.model.DEFAULT_VZA, n)
    assert q.Nquad == o.Nquad == 15 and q.Nstreams == 10
    assert np.array_equal(q.qp_mu, o.qp_mu) and np.array_equal(q.wt_muN, o.wt_muN)
    assert q.imu0 == o.imu0 + 1 and q.imu0Nstart == n * o.imu0 + 1


@pytest.mark.parametrize("pol_n", [1, 3, 4])
@pytest.mark.parametrize("m", [0, 1, 2, 5, 19])
def test_z_moments_match_oracle(pol_n, m):
    q = vb.rt_set_streams(10, 60.0, vb.model.DEFAULT_VZA, pol_n)
    g = vb.hg_greek(0.7, 20)
    a = vb.compute_Z_moments(pol_n, q.qp_mu, g, m)
    b = orc.compute_Z_moments(pol_n, q.qp_mu, orc.GreekCoefs(g.alpha, g.beta, g.gamma, g.delta, g.epsilon, g.zeta), m)
    for x, y in zip(a, b):
        assert np.max(np.abs(x - y)) <= 1e-12 * max(1.0, np.max(np.abs(y)))


def test_z_tables_share_the_legendre_functions_across_scatterers():
    # z_tables evaluates P/R/T once per moment for the longest Greek expansion and slices it per scatterer
    # (Rayleigh: 3 terms, aerosol: 20): same tables as the oracle's per-scatterer evaluation
    from vsmom_b200 import rt_run as rr
    from helpers import to_oracle
    m = vb.synthetic_o2a_model(nSpec=4, nstreams=6, polarization_type="Stokes_IQU", aerosol=True)
    Zpp, Zmp = rr.z_tables(m)
    om = to_oracle(m)
    comps = [om.greek_rayleigh] + [a.greek for a in om.aerosols]
    assert Zpp.shape[:2] == (m.m_max + 1, len(comps))
    for mm in range(m.m_max + 1):
        for c, g in enumerate(comps):
            zp, zm = orc.compute_Z_moments(om.pol_n, om.quad.qp_mu, g, mm)
            assert np.max(np.abs(Zpp[mm, c] - zp.T)) <= 1e-12 * max(1.0, np.abs(zp).max())
            assert np.max(np.abs(Zmp[mm, c] - zm.T)) <= 1e-12 * max(1.0, np.abs(zm).max())


def test_rayleigh_z_has_three_moments():
    q = vb.rt_set_streams(4, 30.0, [0.0], 3)
    zp, zm = vb.compute_Z_moments(3, q.qp_mu, vb.get_greek_rayleigh(0.03), 2)
    assert np.abs(zp).max() > 0


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("aerosol", [False, True])
def test_layer_control_matches_oracle(FT, aerosol):
    m = vb.synthetic_o2a_model(nSpec=200, aerosol=aerosol, float_type=FT)
    c = vb.layer_control(m)
    o = orc.band_control(to_oracle(m), 0)
    assert list(c.ndoubl) == o.ndoubl
    assert list(c.scatter) == [int(s) for s in o.scatter]
    assert list(c.iface) == o.interface
    assert c.ndoubl.max() >= 4          # thick lower layers need several doublings


def test_doubling_number_edges():
    # rt_helper_functions.jl:49-69: tau_end <= dtau_max -> 0 ; exact powers of two hit the eps tie rule
    assert vb.model.doubling_number(1e-3, 5e-4, np.float64) == 0
    assert vb.model.doubling_number(1e-3, 1e-3, np.float64) == 0
    assert vb.model.doubling_number(0.125, 1.0, np.float64) == 3
    assert vb.model.doubling_number(0.1, 1.0, np.float64) == 4
    for nd_ref, (a, b) in {3: (0.125, 1.0), 10: (2.0 ** -10, 1.0)}.items():
        assert orc.doubling_number(a, b)[1] == nd_ref


def test_interface_sequence_with_nonscattering_layers():
    # get_scattering_interface (rt_helper_functions.jl:15-33)
    seq = []
    prev = 0
    for iz, s in enumerate([False, False, True, False, True]):
        prev = orc.get_scattering_interface(prev, s, iz)
        seq.append(prev)
    assert seq == [0, 0, 1, 2, 3]


def test_synthetic_scene_shape():
    m = vb.synthetic_o2a_model(nSpec=100)
    assert m.quad_points.qp_muN.size == 45 and m.Nz == 34 and m.m_max == 19
    col = m.tau_abs.sum(axis=1)
    assert col.max() <= 100.0 * 1.01 and col.min() > 0
    assert abs(m.tau_aer.sum() - 0.3) < 1e-12


def test_rpv_fourier_moments_match_oracle_restatement():
    # host mirror (vectorised) against the oracle's loop restatement of rpv_surface.jl:113-196
    import vsmom_oracle as orc
    from helpers import to_oracle
    m = vb.synthetic_o2a_model(nSpec=2, nstreams=3, polarization_type="Stokes_IQU", aerosol=False)
    got = vb.rpv_surface_rho(m.quad_points, m.pol_n, 3, rho0=0.2, rho_c=0.7, k=0.8, theta_hs=-0.1)
    ref = orc.rpv_surface_rho(to_oracle(m).quad, 3, 0.2, 0.7, 0.8, -0.1)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    # degenerate RPV = Lambertian: 2*albedo in the Stokes-I block of m = 0, nothing else
    lam = vb.rpv_surface_rho(m.quad_points, m.pol_n, 2, rho0=0.3, rho_c=1.0, k=1.0, theta_hs=0.0)
    assert np.allclose(lam[0, ::3, ::3], 0.6, atol=1e-13) and np.abs(lam[1:]).max() < 1e-13
    assert np.abs(lam[0, 1::3]).max() == 0 and np.abs(lam[0, :, 1::3]).max() == 0


def test_prepare_rejects_wrong_surface_rho_shape():
    from vsmom_b200 import rt_run as rr
    m = vb.synthetic_o2a_model(nSpec=2, nstreams=3, polarization_type="Stokes_I", aerosol=False)
    N = m.quad_points.qp_muN.size
    m.surface_rho = np.zeros((m.m_max + 1, N, N + 1))
    with pytest.raises(ValueError):
        rr.prepare(m, "cpu")
    m.surface_rho = np.zeros((m.m_max + 1, N, N))
    p = rr.prepare(m, "cpu")             # host-side preparation needs no GPU
    assert p.surf_rho is not None and tuple(p.surf_rho.shape) == (m.m_max + 1, N, N)
    m.surface_rho = None
    assert rr.prepare(m, "cpu").surf_rho is None


def test_rossli_fourier_moments_match_oracle_restatement():
    # host mirror (vectorised) against the oracle's loop restatement of rossli_surface.jl
    import vsmom_oracle as orc
    from helpers import to_oracle
    m = vb.synthetic_o2a_model(nSpec=2, nstreams=3, polarization_type="Stokes_IQ", aerosol=False)
    got = vb.rossli_surface_rho(m.quad_points, m.pol_n, 3, fiso=0.1, fvol=0.05, fgeo=0.02)
    ref = orc.brdf_surface_rho(to_oracle(m).quad, 3, lambda a, b, p: orc.rossli_reflectance(0.1, 0.05, 0.02, a, b, p))
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    # reciprocity of the kernels: rho(mu_i, mu_r) == rho(mu_r, mu_i) -> symmetric Stokes-I block
    assert np.max(np.abs(got[:, ::2, ::2] - got[:, ::2, ::2].transpose(0, 2, 1))) <= 1e-12
    # isotropic only == Lambertian: 2*fiso at m = 0
    iso = vb.rossli_surface_rho(m.quad_points, m.pol_n, 2, fiso=0.25, fvol=0.0, fgeo=0.0)
    assert np.allclose(iso[0, ::2, ::2], 0.5, atol=1e-13) and np.abs(iso[1:]).max() < 1e-13


# ---- Raman host logic (vsmom_b200/raman.py)
def test_raman_host_helpers():
    from vsmom_b200 import raman
    rs = raman.synthetic_rrs()
    # normalize_raman_weights! (Inelastic/types.jl:625-629): the weights sum to the Raman fraction 1 - ϖ_Cabannes
    assert abs(rs.varpi_raman.sum() - 0.04) < 1e-15 and np.all(rs.varpi_raman > 0)
    assert len(set(rs.i_shift.tolist())) == len(rs.i_shift) and 0 not in rs.i_shift
    assert rs.i_shift.min() < 0 < rs.i_shift.max()                 # Stokes and anti-Stokes branches
    g = raman.get_greek_raman(0.75 * 0.1)                          # inelastic_helper.jl:862-880
    assert g.beta[0] == 1.0 and g.alpha[2] == pytest.approx(3 * (1 - 0.075) / (1 + 0.0375))
    m = raman.synthetic_rrs_model(nSpec=50)
    assert m.quad_points.qp_muN.size == 15 and m.Nz == 5 and m.m_max == 2      # config 5 shape
    zp, zm = raman.raman_z_tables(m, rs)
    assert zp.shape == (3, 15, 15) and zm.shape == (3, 15, 15)


def test_raman_owner_ranges_are_balanced_by_units():
    from vsmom_b200 import raman
    rs = raman.synthetic_rrs()
    n = 6837
    w = raman.rrs_unit_weights(n, rs)
    for world in (1, 2, 3, 8):
        b = raman.rrs_shard_bounds(n, rs, world)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = [w[lo:hi].sum() for lo, hi in b]
        assert max(loads) - min(loads) <= 2 * w.max()              # equal up to one wavelength's worth of units
    # band-edge wavelengths have fewer in-band donors -> the edge ranges are longer
    b8 = raman.rrs_shard_bounds(n, rs, 8)
    assert (b8[0][1] - b8[0][0]) > (b8[3][1] - b8[3][0])
    # degenerate: more ranks than points still covers the band
    b = raman.rrs_shard_bounds(3, rs, 8)
    assert b[0][0] == 0 and b[-1][1] == 3 and all(lo <= hi for lo, hi in b)

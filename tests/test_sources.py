"""Extra sources of the reference's source-term seam: the ThermalEmission per-source slot
(Sources/thermal_emission.jl, doubling.jl:66-80, interaction.jl:231-259) and SurfaceSIF (Sources/surface_sif.jl).
CPU: the oracle against the reference's own known answers (test/test_thermal_emission.jl T-A1, T-A2, T-A7);
GPU: the CUDA path against the oracle."""
import math

import numpy as np
import pytest

import vsmom_b200 as vb
import vsmom_oracle as orc
from helpers import relerr, to_oracle


def _nadir_model(nspec=3, pol="Stokes_I"):
    return vb.synthetic_o2a_model(nSpec=nspec, nstreams=4, polarization_type=pol, aerosol=False, vza=[0.0], vaz=[0.0])


def test_thermal_elemental_formula_T_A1():
    # test_thermal_emission.jl T-A1: 2 pi (1 - varpi) B (1 - exp(-dtau/mu)) on the I rows, Q/U/V rows stay zero
    m = _nadir_model(pol="Stokes_IQU")
    om = to_oracle(m)
    dtau = np.array([1e-3, 5e-3, 2e-2]); varpi = np.array([0.2, 0.5, 0.9]); B = np.array([3.0, 4.0, 5.0])
    j = orc.thermal_elemental(om, dtau, varpi, B)
    mu = om.quad.qp_muN
    for s in range(3):
        for i in range(mu.size):
            want = 2 * math.pi * (1 - varpi[s]) * B[s] * (1 - math.exp(-dtau[s] / mu[i])) if i % 3 == 0 else 0.0
            assert abs(j[s, i] - want) <= 1e-15 * max(1.0, abs(want))


def test_thick_isothermal_column_is_a_blackbody_T_A7():
    # test_thermal_emission.jl T-A7: opaque isothermal column -> R/B = 1 (atol 1e-3) for every forced m_max
    # (the slot only fires at m = 0: T-A2)
    for m_force in (0, 1, 2):
        m = _nadir_model()
        m.tau_abs[:] = 100.0
        m.m_max = m_force
        base = orc.rt_run(to_oracle(m))
        m.thermal_B = np.full((m.nSpec, m.Nz), 7.5)
        out = orc.rt_run(to_oracle(m))
        ratio = (out["R_SFI"][0, 0, 0] - base["R_SFI"][0, 0, 0]) / 7.5
        assert abs(ratio - 1.0) < 1e-3


def test_surface_sif_is_linear_and_unattenuated_at_zero_optical_depth():
    # 2*SIF0 enters the surface j0- at m = 0 (surface_sif.jl:168-178): the TOA signal is linear in SIF0 and, for a
    # vanishing atmosphere, equals SIF0/pi (Lambertian radiance of the emitted irradiance)
    m = _nadir_model()
    m.tau_abs[:] = 1e-12; m.tau_rayl[:] = 1e-9
    base = orc.rt_run(to_oracle(m))
    m.sif0 = np.full((m.nSpec, 1), 2.0)
    one = orc.rt_run(to_oracle(m))
    m.sif0 = np.full((m.nSpec, 1), 6.0)
    three = orc.rt_run(to_oracle(m))
    d1 = one["R_SFI"] - base["R_SFI"]; d3 = three["R_SFI"] - base["R_SFI"]
    assert np.allclose(d3, 3 * d1, rtol=1e-12)
    assert np.allclose(d1[0, 0], 2.0 / math.pi, rtol=1e-6)


# ---------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx():
    from vsmom_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def _with_sources(m, seed=0):
    rng = np.random.default_rng(seed)
    m.thermal_B = 0.5 + rng.random((m.nSpec, m.Nz))
    if m.surface_rho is None:
        sif = np.zeros((m.nSpec, m.pol_n)); sif[:, 0] = 0.1 + rng.random(m.nSpec)
        m.sif0 = sif
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["generic_small", "fast_n45", "stream_n60"])
def test_sources_match_oracle_on_gpu(ctx, kind):
    from vsmom_b200 import rt_run as rr
    if kind == "generic_small":
        m = vb.synthetic_o2a_model(nSpec=6, nstreams=4, polarization_type="Stokes_IQU", aerosol=True)
    elif kind == "fast_n45":
        m = vb.synthetic_o2a_model(nSpec=6, aerosol=True)
    else:
        m = vb.synthetic_ocean_model(nSpec=2, nstreams=11, polarization_type="Stokes_IQUV")
        m.m_max = 3; m.surface_rho = m.surface_rho[:4]
        for a in m.aerosol_optics:
            g = a.greek_coefs
            for f in ("alpha", "beta", "gamma", "delta", "epsilon", "zeta"):
                setattr(g, f, getattr(g, f)[:4])
    m.tau_abs = m.tau_abs * np.logspace(-4.0, -1.0, m.nSpec)[:, None]
    base = rr.rt_run(m, ctx)
    _with_sources(m)
    out = rr.rt_run(m, ctx)
    ref = orc.rt_run(to_oracle(m))
    for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw"):
        assert np.isfinite(out[k]).all(), k
        assert relerr(out[k], ref[k]) <= 1e-9, (k, relerr(out[k], ref[k]))
    assert relerr(out["R_SFI"], base["R_SFI"]) > 1e-3          # the sources are really in


# ---- single-scattering driver (rt_run_ss)
@pytest.mark.gpu
@pytest.mark.parametrize("dtype,pol,brdf", [(np.float64, "Stokes_IQU", False), (np.float64, "Stokes_I", True), (np.float32, "Stokes_IQU", False)])
def test_rt_run_ss_matches_oracle(dtype, pol, brdf):
    import vsmom_oracle as orc
    from helpers import relerr, to_oracle
    from vsmom_b200 import rt_run as rr
    m = vb.synthetic_o2a_model(nSpec=40, nstreams=6, polarization_type=pol, aerosol=True, float_type=dtype)
    if brdf:
        N = m.quad_points.qp_muN.size
        m.surface_rho = 0.1 * np.random.default_rng(0).random((m.m_max + 1, N, N))
    ref = orc.rt_run_ss(to_oracle(m))
    out = rr.rt_run_ss(m)
    tol = 1e-9 if dtype == np.float64 else 1e-3
    for k in ("R_SFI", "T_SFI", "hem_R", "hem_T"):
        assert relerr(out[k], ref[k]) <= tol, (k, relerr(out[k], ref[k]))
    ms = rr.rt_run(m)
    assert np.all(out["R_SFI"][:, 0] <= ms["R_SFI"][:, 0] * (1 + 1e-5) + 1e-10)      # SS <= MS on Stokes I (test_forward_ss.jl:62-73)

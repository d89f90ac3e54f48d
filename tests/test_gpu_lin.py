"""LinMode on the GPU (vsm_rt_column_lin_f64) against the FD-pinned CPU oracle (oracle.rt_run_lin, itself pinned by
finite differences of the forward oracle in tests/test_oracle_lin.py -- the way the reference pins its own LinMode,
test/test_forward_lin.jl:167-244): forward R/T at the forward tolerance, Jacobians to 1e-8 of their largest entry."""
import numpy as np
import pytest

import vsmom_b200 as vb
import vsmom_oracle as orc
from helpers import relerr, to_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from vsmom_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def lin_params(m, seed=0):
    """P = 4 directions: gas column scaling, Rayleigh scaling, aerosol optical depth (tau_ref) scaling, albedo."""
    rng = np.random.default_rng(seed)
    nS, Nz, nA = m.nSpec, m.Nz, len(m.aerosol_optics)
    P = 4
    d_rayl = np.zeros((P, nS, Nz)); d_abs = np.zeros((P, nS, Nz)); d_aer = np.zeros((P, nA, Nz)); d_alb = np.zeros(P)
    d_abs[0] = m.tau_abs * (1.0 + 0.3 * rng.random(Nz))[None, :]
    d_rayl[1] = m.tau_rayl
    if nA:
        d_aer[2] = m.tau_aer
    else:
        d_abs[2] = m.tau_abs * rng.random((nS, Nz))
    d_alb[3] = 1.0
    return d_rayl, d_abs, d_aer, d_alb


def check(m, ctx, tol_fwd=1e-9, tol_jac=1e-8):
    from vsmom_b200 import rt_run as rr
    d_rayl, d_abs, d_aer, d_alb = lin_params(m)
    out = rr.rt_run_lin(m, d_rayl, d_abs, d_aer, d_alb, ctx=ctx)
    ref = orc.rt_run_lin(to_oracle(m), d_rayl, d_abs, d_aer, d_alb)
    for k in ("R_SFI", "T_SFI"):
        assert np.isfinite(out[k]).all()
        assert relerr(out[k], ref[k]) <= tol_fwd, (k, relerr(out[k], ref[k]))
    for k in ("dR_SFI", "dT_SFI"):
        assert np.isfinite(out[k]).all()
        for p in range(out[k].shape[0]):
            scale = max(np.abs(ref[k][p]).max(), 1e-300)
            if np.abs(ref[k][p]).max() == 0:
                assert np.abs(out[k][p]).max() == 0
                continue
            e = np.abs(out[k][p] - ref[k][p]).max() / scale
            assert e <= tol_jac, (k, p, e)
    return out, ref


def test_lin_small_scalar_lambertian(ctx):
    # generic kernels, N = 5 (scalar, 3 streams + 2 nodes), Rayleigh + aerosol + gas, Lambertian albedo parameter
    m = vb.synthetic_o2a_model(nSpec=5, nstreams=3, polarization_type="Stokes_I", aerosol=True, albedo=0.3,
                               p_half=np.array([1.0, 300.0, 700.0, 1000.0]))
    m.tau_abs = m.tau_abs * np.logspace(-4.0, -1.0, 5)[:, None]
    check(m, ctx)


def test_lin_vector_generic(ctx):
    # generic kernels, N = 21 (IQU, 4 streams), 6 layers
    m = vb.synthetic_o2a_model(nSpec=4, nstreams=4, polarization_type="Stokes_IQU", aerosol=True, albedo=0.15,
                               p_half=np.array([1.0, 100.0, 300.0, 500.0, 700.0, 900.0, 1000.0]))
    m.tau_abs = m.tau_abs * np.logspace(-4.0, -1.0, 4)[:, None]
    check(m, ctx)


def test_lin_forward_matches_forward_kernels(ctx):
    # the forward part of the LinMode call equals the forward column
    from vsmom_b200 import rt_run as rr
    m = vb.synthetic_o2a_model(nSpec=6, nstreams=4, polarization_type="Stokes_IQU", aerosol=True)
    m.tau_abs = m.tau_abs * np.logspace(-4.0, -1.0, 6)[:, None]
    d = lin_params(m)
    lin = rr.rt_run_lin(m, *d, ctx=ctx)
    fwd = rr.rt_run(m, ctx)
    for k in ("R_SFI", "T_SFI"):
        assert relerr(lin[k], fwd[k]) <= 1e-11, k


def test_lin_streamed_ocean_n60(ctx):
    # BASELINE.json configs[3] shape (LinMode on the ocean_coxmunk scene), streamed DMMA products: N = 60 (IQUV,
    # 11 streams), 33 layers, Cox-Munk BRDF moments (no surface parameter: the albedo direction has zero Jacobian)
    m = vb.synthetic_ocean_model(nSpec=2, nstreams=11, polarization_type="Stokes_IQUV", ss_correction=False)
    m.m_max = 2
    m.surface_rho = m.surface_rho[:3]
    for a in m.aerosol_optics:
        g = a.greek_coefs
        for f in ("alpha", "beta", "gamma", "delta", "epsilon", "zeta"):
            setattr(g, f, getattr(g, f)[:3])
    assert m.quad_points.qp_muN.size == 60
    out, ref = check(m, ctx)
    assert np.abs(out["dR_SFI"][3]).max() == 0


def test_lin_embedded_n45_runs_on_the_streamed_kernels(ctx):
    # config-2 shape with Jacobians: N = 45 (IQU, 10 streams) has no LinMode tensor-core kernel of its own; it runs
    # zero-padded on the N = 60 streamed kernels (dummy zero-weight streams).  Lambertian albedo parameter included.
    m = vb.synthetic_o2a_model(nSpec=3, nstreams=10, polarization_type="Stokes_IQU", aerosol=True, albedo=0.2,
                               p_half=np.array([1.0, 200.0, 500.0, 800.0, 1000.0]))
    m.tau_abs = m.tau_abs * np.logspace(-3.0, -1.0, 3)[:, None]
    for a in m.aerosol_optics:
        g = a.greek_coefs
        for f in ("alpha", "beta", "gamma", "delta", "epsilon", "zeta"):
            setattr(g, f, getattr(g, f)[:4])
    m.m_max = 3
    assert m.quad_points.qp_muN.size == 45
    check(m, ctx)

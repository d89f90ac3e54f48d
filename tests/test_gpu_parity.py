"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the
same seeded inputs, against the reference's golden tables, and -- at BASELINE.json's
full size -- through size-independent properties.

Tolerances: fp64 GPU vs fp64 oracle  max|Δ|/max|ref| <= 1e-9;
fp32 GPU vs fp64 oracle <= 1e-2 max-rel on I (test/test_float32.jl:64), target 1e-3."""
import math

import numpy as np
import pytest
import torch

import vsmom_b200 as vb
import vsmom_oracle as orc
from helpers import golden, rayleigh_slab_model, relerr, to_oracle
from test_oracle_golden import siewert_compare, siewert_model

pytestmark = pytest.mark.gpu
F64_TOL = 1e-9
KEYS = ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw")


@pytest.fixture(scope="module")
def ctx():
    from vsmom_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def run_gpu(model, ctx, **kw):
    from vsmom_b200 import rt_run as rr
    return rr.rt_run(model, ctx, **kw)


def assert_close(out, ref, tol, keys=KEYS, sel=slice(None)):
    for k in keys:
        a = out[k][..., sel] if out[k].ndim == 3 else out[k][sel]
        assert np.isfinite(a).all(), k
        e = relerr(a, ref[k])
        assert e <= tol, (k, e)


def test_quickstart_scalar_config(ctx):
    # configs[0]: quickstart.yaml shape -- Stokes_I, 3 streams (+sza/vza node), 2 layers, 16 points
    m = vb.synthetic_o2a_model(nSpec=16, nstreams=3, polarization_type="Stokes_I", aerosol=False, vza=[60.0],
                               vaz=[180.0], p_half=np.array([100.0, 500.0, 1000.0]))
    assert m.quad_points.qp_muN.size == 3          # GL node 0.5 == cos(60 deg): unique() merges it with the SZA/VZA node
    assert_close(run_gpu(m, ctx), orc.rt_run(to_oracle(m)), F64_TOL)


@pytest.mark.parametrize("aerosol,nspec,nref", [(False, 96, 96), (True, 40, 12)])
def test_o2a_vector_config_f64(ctx, aerosol, nspec, nref):
    # configs[1]: IQU, 10 streams (N=45), 34 layers
    m = vb.synthetic_o2a_model(nSpec=nspec, aerosol=aerosol)
    out = run_gpu(m, ctx)
    ref = orc.rt_run(to_oracle(m), slice(0, nref))
    assert_close(out, ref, F64_TOL, sel=slice(0, nref))


def thick_o2a_model(nspec, scale):
    """configs[1] shape with the scattering optical depths scaled up so that ||r r|| and ||r R|| of
    the doubling / adding steps span every route of the adaptive inverse (Neumann K=2,4,8 and
    Gauss-Jordan)."""
    m = vb.synthetic_o2a_model(nSpec=nspec, aerosol=True, albedo=0.8)
    m.tau_rayl = m.tau_rayl * scale
    m.tau_aer = m.tau_aer * scale
    # a 12-point grid sits in a line wing that the generator normalises to tau_abs = 100;
    # spread the column absorption over 1e-3 .. 30 instead
    m.tau_abs = m.tau_abs * np.logspace(-5.0, -0.5, nspec)[:, None]
    return m


@pytest.mark.parametrize("scale", [1.0, 8.0, 40.0])
def test_inverse_routes_match_oracle_and_each_other(ctx, scale):
    # vsm_set_inverse_mode: 1 = adaptive Neumann / Gauss-Jordan (default), 0 = Gauss-Jordan only
    m = thick_o2a_model(12, scale)
    ref = orc.rt_run(to_oracle(m))
    try:
        ctx.set_inverse_mode(0)
        gj = run_gpu(m, ctx)
        ctx.set_inverse_mode(1)
        ad = run_gpu(m, ctx)
    finally:
        ctx.set_inverse_mode(1)
    assert_close(gj, ref, F64_TOL)
    assert_close(ad, ref, F64_TOL)
    # the two routes differ by rounding only: on this (albedo 0.8, up to 16 doublings) problem the
    # fp64 LAPACK oracle itself sits 6e-12..3e-11 from an 80-bit evaluation, and so do both routes
    # (oracle/check_inverse_routes_longdouble.py)
    for k in KEYS:
        assert relerr(ad[k], gj[k]) <= 5e-10, k


def test_brdf_surface_matches_oracle(ctx):
    # create_surface_layer!(::AbstractSurfaceType) (rpv_surface.jl:51-97): an RPV surface, every Fourier
    # moment reflects.  configs[1] shape (N=45, specialised kernels) and a small generic-N scene.
    for kw in (dict(nSpec=10, aerosol=True), dict(nSpec=6, nstreams=4, polarization_type="Stokes_IQUV", aerosol=True)):
        m = vb.synthetic_o2a_model(**kw)
        m.tau_abs = m.tau_abs * np.logspace(-5.0, -1.0, m.nSpec)[:, None]
        m.surface_rho = vb.rpv_surface_rho(m.quad_points, m.pol_n, m.m_max, rho0=0.2, rho_c=0.7, k=0.8, theta_hs=-0.1)
        assert_close(run_gpu(m, ctx), orc.rt_run(to_oracle(m)), F64_TOL)


def test_degenerate_rpv_is_lambertian(ctx):
    # rho0 = albedo, k = 1, Theta = 0, rho_c = 1  ->  rho == albedo: the BRDF path must give the Lambertian result
    m = vb.synthetic_o2a_model(nSpec=8, aerosol=True, albedo=0.3)
    m.tau_abs = m.tau_abs * np.logspace(-5.0, -1.0, 8)[:, None]
    lam = run_gpu(m, ctx)
    m.surface_rho = vb.rpv_surface_rho(m.quad_points, m.pol_n, m.m_max, rho0=0.3, rho_c=1.0, k=1.0, theta_hs=0.0)
    brdf = run_gpu(m, ctx)
    # (the m > 0 moments of the BRDF are quadrature noise ~1e-17, not exact zeros.)  T_SFI is left out:
    # create_surface_layer!(::AbstractSurfaceType) injects the direct beam j0+ = I0 exp(-tau/mu0) for every
    # m (rpv_surface.jl:83), the Lambertian one only for m = 0, so T at a VZA that shares the solar stream
    # differs by construction.
    for k in ("R_SFI", "hdr", "bhr_uw", "bhr_dw"):
        assert relerr(brdf[k], lam[k]) <= F64_TOL, k


def test_rt_run_streams_matches_oracle_and_resynthesises(ctx):
    # rt_run_streams (rt_run.jl:107-193): composite R-+, T++, J0-+ per Fourier moment at all streams; and the
    # reference's self-consistency KAT (test/test_CoreRT.jl:45-108): Fourier re-synthesis from the per-m J-
    # equals rt_run, atol 1e-12.  configs[1] shape (specialised kernels) and a generic-N scene.
    from vsmom_b200 import rt_run as rr
    for kw in (dict(nSpec=6, aerosol=True), dict(nSpec=5, nstreams=4, polarization_type="Stokes_IQUV", aerosol=True)):
        m = vb.synthetic_o2a_model(**kw)
        m.tau_abs = m.tau_abs * np.logspace(-5.0, -1.0, m.nSpec)[:, None]
        out = rr.rt_run(m, ctx, return_streams=True)
        ref = orc.rt_run(to_oracle(m), return_streams=True)
        st = out["streams"]
        for mi, C in enumerate(ref["streams"]):
            for got, want, name in ((st["Rmp"][mi], C.Rmp, "Rmp"), (st["Tpp"][mi], C.Tpp, "Tpp"),
                                    (st["Jm"][mi], C.J0m, "Jm"), (st["Jp"][mi], C.J0p, "Jp")):
                assert relerr(got, want) <= F64_TOL, (mi, name, relerr(got, want))
        q, n = m.quad_points, m.pol_n
        R = np.zeros_like(out["R_SFI"])
        for mi in range(m.m_max + 1):
            wgt = (0.5 if mi == 0 else 1.0) / math.pi
            for iv, (vza, vaz) in enumerate(zip(m.vza, m.vaz)):
                imu = orc.nearest_point(q.qp_mu, orc.cosd(vza))
                c, s = orc.cosd(mi * vaz), orc.sind(mi * vaz)
                wv = wgt * np.array([c, c, s, s][:n])
                R[iv] += wv[:, None] * st["Jm"][mi][:, imu * n:(imu + 1) * n].T
        assert np.max(np.abs(R - out["R_SFI"])) <= 1e-12
    s = rr.rt_run_streams(m, ctx)
    assert s.Rmp_per_m.shape == (m.m_max + 1, m.nSpec, q.qp_muN.size, q.qp_muN.size) and s.weight[0] == 0.5 / np.pi


@pytest.mark.parametrize("pol", ["Stokes_I", "Stokes_IQ", "Stokes_IQUV"])
def test_other_polarisations(ctx, pol):
    m = vb.synthetic_o2a_model(nSpec=8, nstreams=4, polarization_type=pol, aerosol=True)
    assert_close(run_gpu(m, ctx), orc.rt_run(to_oracle(m)), F64_TOL)


def test_o2a_vector_config_f32(ctx):
    m32 = vb.synthetic_o2a_model(nSpec=32, aerosol=True, float_type=np.float32)
    m64 = vb.synthetic_o2a_model(nSpec=32, aerosol=True)
    # (a 32-point grid sits in a line wing normalised to tau_abs = 100, where the transmitted field underflows in
    #  fp32: spread the column absorption over 1e-3 .. 30 so that T_SFI is a meaningful comparison)
    sc = np.logspace(-5.0, -0.5, 32)[:, None]
    m32.tau_abs = (m32.tau_abs * sc).astype(np.float32)
    m64.tau_abs = m64.tau_abs * sc
    out = run_gpu(m32, ctx)
    ref64 = orc.rt_run(to_oracle(m64), slice(0, 8))
    I32, I64 = out["R_SFI"][:, 0, :8], ref64["R_SFI"][:, 0]
    assert np.isfinite(out["R_SFI"]).all() and np.isfinite(out["T_SFI"]).all()
    assert np.max(np.abs(I32 - I64) / np.abs(I64)) < 1e-2          # test/test_float32.jl:64
    ref32 = orc.rt_run(to_oracle(m32), slice(0, 8))
    # every output array: vs the fp64 oracle at the reference's F32 bound (1e-2), vs the fp32 oracle at 1e-3
    for k in KEYS:
        a = out[k][..., :8] if out[k].ndim == 3 else out[k][:8]
        assert np.isfinite(out[k]).all(), k
        assert relerr(a, ref64[k]) < 1e-2, (k, relerr(a, ref64[k]))
        assert relerr(a, ref32[k]) < 1e-3, (k, relerr(a, ref32[k]))


def test_natraj_golden_on_gpu(ctx):
    # test/test_CoreRT.jl:110-157 straight against the published table (N = 108, IQUV)
    d = golden()["natraj"]
    vza = [math.degrees(math.acos(x)) for x in d["mu"]]
    for k in (1, 3):
        phi = d["phi"][k]
        m = rayleigh_slab_model("Stokes_IQUV", d["nstreams"], math.degrees(math.acos(d["mu0"])), vza, [phi] * 16,
                                d["tau"], d["albedo"])
        out = run_gpu(m, ctx)
        R = math.pi * out["R_SFI"]
        I, Q, U = (np.array(d[s])[:, k] for s in "IQU")
        assert np.max(np.abs(I - R[:, 0, 0]) / I) < d["tol"]["I"]
        sel = R[:, 2, 0] >= 0.01
        assert np.max(np.abs(U[sel] - R[sel, 2, 0]) / U[sel]) < d["tol"]["U"]
        assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL, keys=("R_SFI", "T_SFI"))


def test_6sv1_golden_on_gpu(ctx):
    # near-coincident streams (sza 0.0001 vs vza 0): the expdiff_neg / L'Hopital branches
    d = golden()["sixsv1"]
    azs, szas, lam, tau, rho = d["cases"][1]
    worst = 0.0
    for si, sza in enumerate(szas[:2]):
        m = rayleigh_slab_model("Stokes_IQUV", d["nstreams"], sza, d["vza"], [azs[0]] * 16, tau, rho)
        out = run_gpu(m, ctx)
        R = math.pi * out["R_SFI"][:, 0, 0] / m.quad_points.mu0
        tr = np.array(d["R_trues"][1][si][0])
        worst = max(worst, float(np.max(np.abs(tr - R) / tr)))
        assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL)
    assert worst < d["tol"]


def test_natraj_cds_tables_on_gpu(ctx):
    # one (albedo, mu0) slice of the CDS subset straight against the published tables and against the oracle:
    # N = 112, generic kernels, reflecting surface, upwelling and downwelling
    from helpers import CDS_TOL, cds_errors, cds_golden, cds_model
    d = cds_golden()
    m = cds_model(d, d["albedo"][1], d["mu0"][1])
    out = run_gpu(m, ctx)
    for st, (ae, re_) in cds_errors(d, out, 1, 1).items():
        assert ae <= CDS_TOL["abs"] and re_ <= CDS_TOL[st], (st, ae, re_)
    assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL)


def test_vlidort_solar_tester_on_gpu(ctx):
    # 23 layers, Rayleigh + gas + HG aerosol, Stokes_I, 16 moments: straight against VLIDORT and against the oracle
    from helpers import solar_tester_errors, solar_tester_model
    st = golden()["solar_tester"]
    m = solar_tester_model(st, st["sza"][0], st["raz"][1])
    out = run_gpu(m, ctx)
    up, dn = solar_tester_errors(st, out, 0, 1)
    assert up < st["rtol"] and dn < st["rtol"], (up, dn)
    assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL)


def test_vlidort_solar_tester_vector_on_gpu(ctx):
    # polarized 23-layer case with the Problem III Mie aerosol (N = 36): against VLIDORT and against the oracle
    from helpers import solar_tester_vector_errors, solar_tester_vector_model
    G = golden()
    st, sv = G["solar_tester"], G["solar_tester_vector"]
    m = solar_tester_vector_model(st, sv, sv["sza"][0], sv["raz"][0])
    out = run_gpu(m, ctx)
    up, dn = solar_tester_vector_errors(sv, out, 0, 0)
    assert up < sv["rtol_up"] and dn < sv["rtol_dn"], (up, dn)
    assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL)


def test_siewert_golden_on_gpu(ctx):
    # aerosol-only slab with supplied Greek coefficients: Rayleigh weight all-zero branch (types.jl:1279)
    m = siewert_model(90.0)
    out = run_gpu(m, ctx)
    for st, e in siewert_compare(out["R_SFI"], 90.0).items():
        assert e < (1e-2 if st == 1 else 5e-4), (st, e)
    assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL, keys=("R_SFI", "T_SFI"))


def test_nonscattering_layers_all_interfaces(ctx):
    # layers that only absorb exercise ScatteringInterface_00/_01/_10 and zero_added_noscat!
    m = vb.synthetic_o2a_model(nSpec=6, nstreams=4, aerosol=False, p_half=np.array([1., 50., 200., 400., 600., 800., 1000.]))
    m.tau_rayl[:, 0] = 0; m.tau_rayl[:, 1] = 0; m.tau_rayl[:, 3] = 0
    m.tau_abs[:, 0] += 0.05; m.tau_abs[:, 3] += 0.2
    c = vb.layer_control(m)
    assert list(c.iface) == [0, 0, 1, 2, 3, 3]
    assert_close(run_gpu(m, ctx), orc.rt_run(to_oracle(m)), F64_TOL)


def test_nonscattering_layers_fresh_context():
    # the added-layer hand-off is sized from the SCATTERING layers and indexed by their compact position
    # (sched_idx); a fresh context (grow-only buffers still empty) as first call catches any extent mismatch
    from vsmom_b200 import _lib
    m = vb.synthetic_o2a_model(nSpec=6, nstreams=8, polarization_type="Stokes_IQU", aerosol=False,
                               p_half=np.array([1., 50., 200., 400., 600., 800., 1000.]))
    m.tau_rayl[:, 0] = 0; m.tau_rayl[:, 1] = 0; m.tau_rayl[:, 3] = 0
    m.tau_abs[:, 0] += 0.05; m.tau_abs[:, 3] += 0.2
    assert list(vb.layer_control(m).scatter) == [0, 0, 1, 0, 1, 1]
    c = _lib.Context(0)
    try:
        out = run_gpu(m, c)
    finally:
        c.close()
    assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL)


def test_nonfinite_output_counter(ctx):
    # failure detection across the ABI: a healthy column reports 0; a NaN optical depth poisons exactly its point
    m = vb.synthetic_o2a_model(nSpec=8, nstreams=4, aerosol=True)
    run_gpu(m, ctx)
    assert ctx.last_column_nonfinite() == 0
    from vsmom_b200 import rt_run as rr
    prep = rr.prepare(m)                      # control flow from the healthy band
    m.tau_abs[3, 2] = np.nan
    out = run_gpu(m, ctx, prep=prep)
    n_out = out["R_SFI"].shape[0] * out["R_SFI"].shape[1]
    assert ctx.last_column_nonfinite() == n_out
    assert np.isfinite(np.delete(out["R_SFI"], 3, axis=2)).all() and not np.isfinite(out["R_SFI"][..., 3]).any()


def test_single_point_and_empty(ctx):
    m = vb.synthetic_o2a_model(nSpec=1, nstreams=4, aerosol=True)
    assert_close(run_gpu(m, ctx), orc.rt_run(to_oracle(m)), F64_TOL)      # batch == 1 must work
    from vsmom_b200 import rt_run as rr
    prep = rr.prepare(m)
    z = torch.zeros((m.Nz, 0), dtype=torch.float64, device="cuda")
    R, T, H, uw, dw = rr.rt_column_device(ctx, prep, z, z, torch.zeros((0, 3), dtype=torch.float64, device="cuda"))
    assert R.shape[0] == 0


def test_workspace_chunking_is_bit_identical(ctx):
    m = vb.synthetic_o2a_model(nSpec=64, nstreams=4, aerosol=True)
    a = run_gpu(m, ctx)
    ctx.set_workspace_limit(3 << 20)       # forces several spectral chunks / moment batches
    try:
        b = run_gpu(m, ctx)
    finally:
        ctx.set_workspace_limit(0)
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k


def test_shard_invariance_and_determinism(ctx):
    # two half-band shards with full-band control flow == the full band, bit for bit
    from vsmom_b200 import rt_run as rr
    m = vb.synthetic_o2a_model(nSpec=64, aerosol=False)
    prep = rr.prepare(m)
    full = run_gpu(m, ctx, prep=prep)
    again = run_gpu(m, ctx, prep=prep)
    lo = run_gpu(m, ctx, prep=prep, sel=slice(0, 32))
    hi = run_gpu(m, ctx, prep=prep, sel=slice(32, 64))
    for k in ("R_SFI", "T_SFI", "hdr"):
        assert np.array_equal(full[k], again[k])
        assert np.array_equal(full[k], np.concatenate([lo[k], hi[k]], axis=2))


def test_full_size_properties(ctx):
    # BASELINE.json configs[1] size (10k points, Rayleigh-only so it stays short): invariants that
    # need no full-size oracle run -- finite, I > 0, |Q|,|U| <= I (test/test_forward_noRS.jl), spot
    # points equal the oracle, and the strongly absorbing points are darker than the continuum.
    m = vb.synthetic_o2a_model(nSpec=10000, aerosol=False)
    out = run_gpu(m, ctx)
    R = out["R_SFI"]
    assert np.isfinite(R).all() and (R[:, 0] > 0).all()
    assert (np.abs(R[:, 1]) <= R[:, 0] + 1e-15).all() and (np.abs(R[:, 2]) <= R[:, 0] + 1e-15).all()
    idx = np.array([0, 1234, 5000, 9999])
    ref = orc.rt_run(to_oracle(m), idx)
    assert relerr(R[..., idx], ref["R_SFI"]) <= F64_TOL
    col = m.tau_abs.sum(axis=1)
    assert R[4, 0, np.argmax(col)] < R[4, 0, np.argmin(col)]


# ---- L0: batched_mul / batch_inv! / batch_solve!  (KAT recipe: test/test_batched_kernels.jl:3-37)
def _dd(batch, n, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.rand((batch, n, n), generator=g, dtype=torch.float64)
    A = A + n * torch.eye(n, dtype=torch.float64)
    return A.to(dtype).cuda()


@pytest.mark.parametrize("n,batch", [(4, 6), (15, 33), (45, 200), (112, 7), (200, 3), (45, 1), (45, 3001)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_batch_inv_matches_lapack(ctx, n, batch, dtype):
    from vsmom_b200 import batched
    A = _dd(batch, n, dtype)
    ref = torch.linalg.inv(A.double().cpu())
    X = torch.empty_like(A)
    info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    batched.batch_inv_(ctx, X, A.clone(), info)
    ctx.synchronize()
    tol = 50 * (1.2e-7 if dtype == torch.float32 else 2.3e-16) * n
    assert (X.double().cpu() - ref).abs().max() <= tol * ref.abs().max()
    assert (info == 0).all()


@pytest.mark.parametrize("m,k,nc,batch", [(4, 4, 4, 6), (45, 45, 45, 100), (45, 45, 45, 3001), (45, 45, 1, 100), (112, 112, 112, 5),
                                           (200, 200, 200, 2), (45, 45, 45, 1), (7, 13, 5, 9)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_batched_mul_matches_matmul(ctx, m, k, nc, batch, dtype):
    from vsmom_b200 import batched
    g = torch.Generator().manual_seed(1)
    A = torch.randn((batch, k, m), generator=g, dtype=torch.float64).to(dtype).cuda()   # Julia (m,k,batch)
    B = torch.randn((batch, nc, k), generator=g, dtype=torch.float64).to(dtype).cuda()  # Julia (k,nc,batch)
    Cm = batched.batched_mul(ctx, A, B)
    ctx.synchronize()
    ref = torch.matmul(B.double(), A.double())          # (C^T = B^T A^T in this memory order)
    tol = 50 * (1.2e-7 if dtype == torch.float32 else 2.3e-16) * k
    assert (Cm.double() - ref).abs().max() <= tol * ref.abs().max()


def test_batch_inv_flags_singular(ctx):
    from vsmom_b200 import batched
    A = _dd(3, 8, torch.float64)
    A[1] = 0
    info = torch.zeros(3, dtype=torch.int32, device="cuda")
    batched.batch_inv_(ctx, torch.empty_like(A), A, info)
    ctx.synchronize()
    assert info[0] == 0 and info[2] == 0 and info[1] > 0


def test_batch_solve(ctx):
    from vsmom_b200 import batched
    A = _dd(20, 45, torch.float64)
    g = torch.Generator().manual_seed(2)
    B = torch.randn((20, 3, 45), generator=g, dtype=torch.float64).cuda()
    X = torch.empty_like(B)
    batched.batch_solve_(ctx, X, A.clone(), B)
    ctx.synchronize()
    # memory order: X_torch[k] = X_julia^T ; A_julia X_julia = B_julia  <=>  X_t A_t = B_t
    assert (torch.matmul(X, A) - B).abs().max() < 1e-10


def test_invalid_arguments_fail_loudly(ctx):
    from vsmom_b200 import _lib, batched
    with pytest.raises(_lib.VsmError):
        ctx.check(_lib.lib.vsm_batch_inv_f64(ctx.handle, None, None, 4, 1, None))
    with pytest.raises(ValueError):
        batched.batched_mul(ctx, torch.zeros((2, 3, 3)), torch.zeros((2, 3, 3)))


# ---- L1: vsm_rt_layer (drop-in for rt_kernel!) driven layer by layer like rt_run's loop
@pytest.mark.parametrize("dtype,pol,nstreams", [(np.float64, "Stokes_IQU", 10), (np.float64, "Stokes_IQUV", 4),
                                                (np.float32, "Stokes_IQU", 4)])
def test_rt_layer_entry_matches_oracle(ctx, dtype, pol, nstreams):
    from vsmom_b200 import layers
    m = vb.synthetic_o2a_model(nSpec=6, nstreams=nstreams, polarization_type=pol, aerosol=True, float_type=dtype)
    m.tau_rayl[:, 1] = 0                      # one purely absorbing layer: 10/11 interfaces + zero_added_noscat!
    m.tau_aer[:, 1] = 0
    om = to_oracle(m)
    q = m.quad_points
    N = q.qp_muN.size
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    dev = "cuda"
    mu = torch.from_numpy(q.qp_muN.astype(dtype)).to(dev)
    wt = torch.from_numpy(q.wt_muN.astype(dtype)).to(dev)
    F0 = torch.zeros((m.nSpec, m.pol_n), dtype=tdt, device=dev)
    F0[:, 0] = 1
    for mom in (0, 2):
        ctl = orc.band_control(om, mom)
        props = orc.construct_core_optical_properties(om, mom)
        comp = layers.make_composite_layer(N, m.nSpec, tdt)
        added = layers.make_added_layer(N, m.nSpec, tdt)
        Cref = None
        F0o = np.zeros((m.pol_n, m.nSpec), dtype); F0o[0] = 1
        for iz in range(m.Nz):
            p = props[iz]
            Zpp = np.broadcast_to(p.Zpp, (m.nSpec, N, N)) if p.Zpp.ndim == 2 else p.Zpp
            Zmp = np.broadcast_to(p.Zmp, (m.nSpec, N, N)) if p.Zmp.ndim == 2 else p.Zmp
            tz = lambda Z: torch.from_numpy(np.ascontiguousarray(Z.transpose(0, 2, 1)).astype(dtype)).to(dev)
            layers.rt_kernel_(ctx, m.pol_n, added if iz % 2 == 0 else None, comp,
                              torch.from_numpy(np.ascontiguousarray(p.tau.astype(dtype))).to(dev),
                              torch.from_numpy(np.ascontiguousarray(np.broadcast_to(p.varpi, p.tau.shape).astype(dtype))).to(dev),
                              tz(Zpp), tz(Zmp), ctl.interface[iz],
                              torch.from_numpy(np.ascontiguousarray(ctl.tau_sum[:, iz])).to(dev), mom, q, F0, iz + 1,
                              ctl.ndoubl[iz], ctl.scatter[iz], mu, wt)
            Cref = orc.rt_kernel_layer(om, mom, iz, ctl, p, Cref, F0o, ctl.tau_sum[:, iz])
        ctx.synchronize()
        tol = 1e-9 if dtype == np.float64 else 2e-3
        for name, ref in (("Rmp", Cref.Rmp), ("Rpm", Cref.Rpm), ("Tpp", Cref.Tpp), ("Tmm", Cref.Tmm)):
            got = getattr(comp, name).cpu().numpy().transpose(0, 2, 1)
            assert relerr(got, ref) <= tol, (mom, name, relerr(got, ref))
        assert relerr(comp.J0m.cpu().numpy(), Cref.J0m) <= tol
        assert relerr(comp.J0p.cpu().numpy(), Cref.J0p) <= tol
        # AddedLayer outputs of the last even layer obey the D-symmetry r⁺⁻ = D r⁻⁺ D
        Dd = np.tile(np.array([1, 1, -1, -1][:m.pol_n]), q.Nquad)
        rmp = added.rmp.cpu().numpy().transpose(0, 2, 1)
        rpm = added.rpm.cpu().numpy().transpose(0, 2, 1)
        if np.abs(rmp).max() > 0:
            assert np.allclose(rpm, Dd[None, :, None] * rmp * Dd[None, None, :], rtol=0, atol=0)


# ----------------------------------------------------------------------------
# BASELINE.json configs[2]: ocean_coxmunk.yaml scene (polarized Cox-Munk glint + aerosol), N = 60 / 75 / 100:
# the streamed-operand kernels (vsm_stream.cuh)
# ----------------------------------------------------------------------------
def trim_moments(m, mmax):
    """Shorten the Fourier loop (keeps the oracle run short): m_max, the surface moments and the aerosol
    Greek coefficients are cut consistently for both sides."""
    m.m_max = mmax
    m.surface_rho = m.surface_rho[:mmax + 1]
    for a in m.aerosol_optics:
        g = a.greek_coefs
        for f in ("alpha", "beta", "gamma", "delta", "epsilon", "zeta"):
            setattr(g, f, getattr(g, f)[:mmax + 1])
    return m


@pytest.mark.parametrize("pol,nstreams,N,mmax", [("Stokes_IQU", 20, 75, None), ("Stokes_IQUV", 20, 100, 7),
                                                  ("Stokes_IQUV", 11, 60, 9)])
def test_ocean_coxmunk_config3_matches_oracle(ctx, pol, nstreams, N, mmax):
    m = vb.synthetic_ocean_model(nSpec=2, nstreams=nstreams, polarization_type=pol)
    assert m.quad_points.qp_muN.size == N
    if mmax is not None:
        trim_moments(m, mmax)
    else:
        assert m.m_max == 2 * nstreams - 1           # the full 40-moment Fourier loop of the benchmark row
    out = run_gpu(m, ctx)
    assert_close(out, orc.rt_run(to_oracle(m)), F64_TOL)
    # the TMS correction is in: switching it off changes R_SFI by exactly mu0 exp(-tau/mu0) * factor
    ss = m.ss_correction
    m.ss_correction = None
    off = run_gpu(m, ctx)
    tau = (m.tau_rayl + m.tau_abs).sum(axis=1) + m.tau_aer.sum()
    mu0 = m.quad_points.mu0
    want = ss[:, :, None] * (mu0 * np.exp(-tau / mu0))[None, None, :]
    assert np.allclose(out["R_SFI"] - off["R_SFI"], want, rtol=1e-8, atol=1e-15)


def test_ocean_streamed_equals_generic_many_points(ctx):
    # several work items per CTA (slab reuse), every CTA busy: streamed kernels == the generic shared-memory/slab
    # kernels to rounding on 600 points, and spot points equal the oracle
    import os
    from vsmom_b200 import _lib
    m = trim_moments(vb.synthetic_ocean_model(nSpec=600, nstreams=20, polarization_type="Stokes_IQU"), 2)
    out = run_gpu(m, ctx)
    os.environ["VSM_DISABLE_STREAM"] = "1"
    try:
        c = _lib.Context(0)
    finally:
        del os.environ["VSM_DISABLE_STREAM"]
    try:
        gen = run_gpu(m, c)
    finally:
        c.close()
    for k in KEYS:
        assert relerr(out[k], gen[k]) <= 1e-10, k
    idx = np.array([0, 233, 599])
    ref = orc.rt_run(to_oracle(m), idx)
    for k in ("R_SFI", "T_SFI", "hdr"):
        assert relerr(out[k][..., idx], ref[k]) <= F64_TOL, k


@pytest.mark.parametrize("scale", [8.0, 40.0])
def test_ocean_streamed_inverse_routes(ctx, scale):
    # thick, bright scene: Neumann orders 2..8 and the staged partial-pivoting elimination all occur; the
    # elimination-only mode must agree
    m = trim_moments(vb.synthetic_ocean_model(nSpec=6, nstreams=20, polarization_type="Stokes_IQU"), 3)
    m.tau_rayl = m.tau_rayl * scale
    m.tau_aer = m.tau_aer * scale
    m.tau_abs = m.tau_abs * np.logspace(-3.0, 0.0, 6)[:, None]
    ref = orc.rt_run(to_oracle(m))
    try:
        ctx.set_inverse_mode(0)
        gj = run_gpu(m, ctx)
    finally:
        ctx.set_inverse_mode(1)
    ad = run_gpu(m, ctx)
    assert_close(gj, ref, F64_TOL)
    assert_close(ad, ref, F64_TOL)


# ----------------------------------------------------------------------------
# several GPUs from one process behind the C-ABI (vsm_multi_create / vsm_rt_column_sharded_*)
# ----------------------------------------------------------------------------
def _multi_equals_single(devices, ctx, nspec=130):
    from vsmom_b200 import _lib, rt_run as rr
    m = vb.synthetic_o2a_model(nSpec=nspec, nstreams=4, aerosol=True)
    single = run_gpu(m, ctx)
    mc = _lib.MultiContext(devices)
    try:
        multi = rr.rt_run_multi(m, mc)
        again = rr.rt_run_multi(m, mc)
    finally:
        mc.close()
    for k in KEYS:
        assert np.array_equal(multi[k], single[k]), k          # sharding changes nothing, bit for bit
        assert np.array_equal(multi[k], again[k]), k


def test_multi_context_one_device(ctx):
    # the in-library sharded entry with a single device: staging, NCCL communicator of one rank, packing, unpacking
    _multi_equals_single([0], ctx)


def test_multi_context_two_devices_one_process(ctx):
    # two GPUs driven from this one process (the Julia situation): ragged shards (130 points -> 96 + 34)
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _multi_equals_single([0, 1], ctx)
    m = vb.synthetic_ocean_model(nSpec=40, nstreams=11, polarization_type="Stokes_IQUV")
    trim_moments(m, 2)
    from vsmom_b200 import _lib, rt_run as rr
    mc = _lib.MultiContext([0, 1])
    try:
        multi = rr.rt_run_multi(m, mc)
    finally:
        mc.close()
    single = run_gpu(m, ctx)
    for k in KEYS:
        assert np.array_equal(multi[k], single[k]), k


def test_small_matrix_kernel_equals_generic_kernels():
    """N <= 16 in fp64 runs the warp-per-point kernel (vsm_small.cuh); VSM_DISABLE_SMALL=1 forces the CTA-per-item
    generic kernels.  Both must agree to rounding on every output, including non-scattering layers and a BRDF surface."""
    import os
    from vsmom_b200 import _lib, rt_run as rr
    rng = np.random.default_rng(4)
    q = vb.rt_set_streams(3, 40.0, [25.0], 3, np.float64)
    Nz, nS = 6, 37
    tau_rayl = 0.02 * rng.random((nS, Nz)) + 0.01
    tau_rayl[:, [0, 3]] = 0.0                        # non-scattering layers -> interfaces 00, 01, 10, 11
    tau_abs = 0.3 * rng.random((nS, Nz))
    N = q.qp_muN.size
    rho = 0.05 * rng.random((3, N, N))
    m = vb.RTModel("Stokes_IQU", q, 40.0, [25.0], [70.0], tau_rayl, tau_abs, np.zeros((0, Nz)), [], vb.get_greek_rayleigh(0.03),
                   0.2, 2, 1.0, vb.RTNumericalParameters(), np.float64)
    for surf in (None, rho):
        m.surface_rho = surf
        res = []
        for dis in (None, "1"):
            if dis:
                os.environ["VSM_DISABLE_SMALL"] = dis
            else:
                os.environ.pop("VSM_DISABLE_SMALL", None)
            c = _lib.Context(0)
            res.append(rr.rt_run(m, ctx=c))
            c.close()
        os.environ.pop("VSM_DISABLE_SMALL", None)
        for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw"):
            assert relerr(res[0][k], res[1][k]) <= 1e-12, (k, relerr(res[0][k], res[1][k]))
        ref = orc.rt_run(to_oracle(m))
        for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw"):
            assert relerr(res[0][k], ref[k]) <= 1e-9, (k, relerr(res[0][k], ref[k]))


@pytest.mark.parametrize("pol,nstreams,expect_N", [("Stokes_I", 12, 17), ("Stokes_I", 16, 21), ("Stokes_IQU", 8, 39), ("Stokes_IQU", 4, 27), ("Stokes_I", 30, 35),
                                                   ("Stokes_IQUV", 8, 52), ("Stokes_IQUV", 12, 68), ("Stokes_IQU", 23, 81)])
def test_embedded_sizes_match_oracle_and_generic(pol, nstreams, expect_N):
    """N between the specialised sizes runs zero-padded as the next one (45 / 60 / 75 / 100) with dummy zero-weight
    streams; VSM_DISABLE_EMBED=1 keeps the run-time-N generic kernels.  Same results, and both match the oracle."""
    import os
    from vsmom_b200 import _lib, rt_run as rr
    m = vb.synthetic_o2a_model(nSpec=5, nstreams=nstreams, polarization_type=pol, aerosol=True)
    for a in m.aerosol_optics:                       # short Fourier loop: keeps the oracle run short
        g = a.greek_coefs
        for f in ("alpha", "beta", "gamma", "delta", "epsilon", "zeta"):
            setattr(g, f, getattr(g, f)[:5])
    m.m_max = 4
    m.tau_rayl[:, 2] = 0.0
    m.tau_aer[:, 2] = 0.0                            # one non-scattering layer
    assert m.quad_points.qp_muN.size == expect_N
    res = []
    for dis in (None, "1"):
        if dis:
            os.environ["VSM_DISABLE_EMBED"] = dis
        else:
            os.environ.pop("VSM_DISABLE_EMBED", None)
        c = _lib.Context(0)
        res.append(rr.rt_run(m, ctx=c))
        c.close()
    os.environ.pop("VSM_DISABLE_EMBED", None)
    ref = orc.rt_run(to_oracle(m))
    for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw"):
        # (two kernel families: the deep-absorption points of this band sit at 1e-50, where exp() arguments of ~115 make
        #  1e-11 relative differences between any two evaluation orders)
        assert relerr(res[0][k], res[1][k]) <= 1e-10, (k, relerr(res[0][k], res[1][k]))
        assert relerr(res[0][k], ref[k]) <= 1e-9, (k, relerr(res[0][k], ref[k]))


def test_embedded_size_fp32():
    """fp32: N = 39 runs embedded in the compile-time N = 45 kernels; <= 1e-2 vs the fp64 oracle (test_float32.jl:64) and
    close to the fp32 generic kernels."""
    import os
    from vsmom_b200 import _lib, rt_run as rr
    m = vb.synthetic_o2a_model(nSpec=6, nstreams=8, polarization_type="Stokes_IQU", aerosol=True, float_type=np.float32)
    assert m.quad_points.qp_muN.size == 39
    sc = np.logspace(-5.0, -1.0, 6)[:, None]         # (as in test_o2a_vector_config_f32: keep T_SFI above the fp32 underflow)
    m.tau_abs = (m.tau_abs * sc).astype(np.float32)
    res = []
    for dis in (None, "1"):
        if dis:
            os.environ["VSM_DISABLE_EMBED"] = dis
        else:
            os.environ.pop("VSM_DISABLE_EMBED", None)
        c = _lib.Context(0)
        res.append(rr.rt_run(m, ctx=c))
        c.close()
    os.environ.pop("VSM_DISABLE_EMBED", None)
    m64 = vb.synthetic_o2a_model(nSpec=6, nstreams=8, polarization_type="Stokes_IQU", aerosol=True)
    m64.tau_abs = m64.tau_abs * sc
    ref = orc.rt_run(to_oracle(m64))
    for k in ("R_SFI", "T_SFI"):
        assert relerr(res[0][k], ref[k]) <= 1e-2, (k, relerr(res[0][k], ref[k]))
        assert relerr(res[0][k], res[1][k]) <= 2e-3, (k, relerr(res[0][k], res[1][k]))

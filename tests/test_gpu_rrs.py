"""GPU parity of the rotational-Raman path (vsm_rt_column_rrs_f64, BASELINE config 5) against the closure-pinned
oracle (oracle/vsmom_oracle_rrs.py), through the C-ABI.  fp64 tolerance: max|Δ| / max|ref| <= 1e-9 per output."""
import os

import numpy as np
import pytest

import vsmom_b200 as vb
from vsmom_b200 import raman
from helpers import relerr, to_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-9
KEYS = ("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI")


def _scene(pol="Stokes_IQU", nstreams=3, nS=40, Nz=3, aerosol=False, albedo=0.2, seed=5, vza=(20.0,), tau=0.05,
           eps=0.04):
    rng = np.random.default_rng(seed)
    q = vb.rt_set_streams(nstreams, 35.0, list(vza), vb.model.POL_N[pol], np.float64)
    tau_rayl = tau * np.linspace(0.6, 1.4, Nz)[None, :] * (1.0 + 0.02 * np.linspace(-1, 1, nS))[:, None]
    lines = np.zeros(nS)
    for c in rng.integers(0, nS, 5):
        lines += 2.0 * np.exp(-0.5 * ((np.arange(nS) - c) / 1.3) ** 2)
    tau_abs = lines[:, None] * np.linspace(0.05, 0.3, Nz)[None, :] + 1e-4
    aer, tau_aer, m_max = [], np.zeros((0, Nz)), 2
    if aerosol:
        aer = [vb.AerosolOptics(vb.hg_greek(0.6, 6), 0.92, 0.0)]
        tau_aer = np.array([np.linspace(0.0, 0.08, Nz)])
        m_max = min(5, 2 * nstreams - 1)
    m = vb.RTModel(pol, q, 35.0, list(vza), [30.0, 140.0][:len(vza)], tau_rayl, tau_abs, tau_aer, aer,
                   vb.get_greek_rayleigh(0.007), albedo, m_max, 1.0 - eps, vb.RTNumericalParameters(), np.float64)
    shifts = np.array([-9, -6, -4, -1, 2, 3, 7, 11], np.int32)
    rs = raman.RRS(shifts, raman.normalize_raman_weights(rng.random(8) + 0.2, 1.0 - eps), raman.get_greek_raman(0.1))
    return m, rs


def _oracle(m, rs, F0=None):
    import vsmom_oracle as orc
    import vsmom_oracle_rrs as orr
    om = to_oracle(m)
    om.F0 = F0
    g = rs.greek_raman
    og = orc.GreekCoefs(*(np.asarray(x, float) for x in (g.alpha, g.beta, g.gamma, g.delta, g.epsilon, g.zeta)))
    return orr.rt_run_rrs(om, orr.RRS(np.asarray(rs.i_shift), np.asarray(rs.varpi_raman), og))


@pytest.mark.parametrize("pol,aerosol", [("Stokes_IQU", False), ("Stokes_IQU", True), ("Stokes_I", False), ("Stokes_IQUV", False)])
def test_rrs_matches_oracle_small_N(pol, aerosol):
    m, rs = _scene(pol=pol, aerosol=aerosol, nstreams=3 if pol != "Stokes_IQUV" else 2)
    assert m.quad_points.qp_muN.size <= 16
    ref = _oracle(m, rs)
    out = raman.rt_run_rrs(m, rs)
    for k in KEYS:
        assert relerr(out[k], ref[k]) <= TOL, (k, relerr(out[k], ref[k]))
    assert np.max(np.abs(ref["ieR_SFI"])) > 1e-4 * np.max(np.abs(ref["R_SFI"]))     # not vacuous


def test_rrs_matches_oracle_structured_solar_spectrum_and_dark_surface():
    m, rs = _scene(albedo=0.0, seed=9)
    F0 = np.zeros((3, m.nSpec))
    F0[0] = 1.0 + 0.4 * np.random.default_rng(1).random(m.nSpec)
    ref = _oracle(m, rs, F0)
    out = raman.rt_run_rrs(m, rs, F0=F0)
    for k in KEYS:
        assert relerr(out[k], ref[k]) <= TOL, (k, relerr(out[k], ref[k]))


@pytest.mark.parametrize("nstreams,lo,hi", [(8, 16, 32), (12, 32, 48)])
def test_rrs_matches_oracle_wide_kernels(nstreams, lo, hi):
    """8 Gauss streams + VZA + SZA -> 10 streams, IQU -> N = 30: the NP = 32 instantiation; 12 streams -> N = 42: NP = 48."""
    m, rs = _scene(nstreams=nstreams, nS=24, Nz=2, aerosol=True)
    assert lo < m.quad_points.qp_muN.size <= hi
    ref = _oracle(m, rs)
    out = raman.rt_run_rrs(m, rs)
    for k in KEYS:
        assert relerr(out[k], ref[k]) <= TOL, (k, relerr(out[k], ref[k]))


def test_rrs_elastic_part_equals_noRS_kernels():
    """The elastic half of the Raman path is an independent (warp-per-point) implementation of the same column the
    noRS kernels compute with ϖ_Cabannes: the two CUDA paths must agree to rounding."""
    from vsmom_b200 import rt_run as rr
    m, rs = _scene(aerosol=True)
    a = raman.rt_run_rrs(m, rs)
    b = rr.rt_run(m)
    assert relerr(a["R_SFI"], b["R_SFI"]) <= 1e-12
    assert relerr(a["T_SFI"], b["T_SFI"]) <= 1e-12


def test_rrs_owner_range_is_bit_identical_to_full_band():
    """Sharding contract: computing only an owner range (with its locally recomputed elastic halo) reproduces the
    corresponding slice of the full-band result bit for bit."""
    m, rs = _scene(nS=48)
    full = raman.rt_run_rrs(m, rs)
    for lo, hi in ((0, 17), (17, 40), (40, 48), (5, 6)):
        part = raman.rt_run_rrs(m, rs, owner=(lo, hi))
        for k in KEYS:
            assert part[k].shape[2] == hi - lo
            assert np.array_equal(part[k], full[k][:, :, lo:hi]), (k, lo, hi)


def test_rrs_unit_enumeration_orders_agree():
    from vsmom_b200 import _lib
    m, rs = _scene(nS=32)
    res = []
    for order in ("0", "1"):
        os.environ["VSM_RRS_ORDER"] = order
        ctx = _lib.Context(0)
        res.append(raman.rt_run_rrs(m, rs, ctx=ctx))
        ctx.close()
    os.environ.pop("VSM_RRS_ORDER")
    for k in KEYS:
        assert np.array_equal(res[0][k], res[1][k])


def test_rrs_rejects_non_scattering_layer():
    from vsmom_b200 import _lib
    m, rs = _scene()
    m.tau_rayl[:, 1] = 0.0
    with pytest.raises(_lib.VsmError):
        raman.rt_run_rrs(m, rs)


def test_rrs_config5_shape_invariants():
    """test/test_forward_raman.jl:90-188 on the config-5 shaped synthetic scene (N = 15, 5 layers), 600 points."""
    m = raman.synthetic_rrs_model(nSpec=600)
    rs = raman.synthetic_rrs(grid_step=2.0)          # the line table compressed onto the short band: shifts within +-100 points
    out = raman.rt_run_rrs(m, rs)
    for k in KEYS:
        assert np.all(np.isfinite(out[k]))
    I_tot = out["R_SFI"][:, 0] + out["ieR_SFI"][:, 0]
    assert np.all(out["R_SFI"][:, 0] > 0) and np.all(I_tot > 0)
    for s in (1, 2):
        assert np.all(np.abs(out["R_SFI"][:, s] + out["ieR_SFI"][:, s]) <= I_tot + 1e-10)
    import copy
    ms = copy.copy(m)
    ms.tau_rayl, ms.tau_abs = m.tau_rayl[0:300], m.tau_abs[0:300]
    ref = _oracle(ms, rs)
    part = raman.rt_run_rrs(ms, rs)
    for k in KEYS:
        assert relerr(part[k], ref[k]) <= TOL, (k, relerr(part[k], ref[k]))


def test_rrs_with_brdf_surface_moments():
    """A non-Lambertian surface (Fourier moments of R_surf for every m) under the Raman column."""
    m, rs = _scene(aerosol=True, seed=21)
    N = m.quad_points.qp_muN.size
    m.surface_rho = 0.08 * np.random.default_rng(8).random((m.m_max + 1, N, N))
    ref = _oracle(m, rs)
    out = raman.rt_run_rrs(m, rs)
    for k in KEYS:
        assert relerr(out[k], ref[k]) <= TOL, (k, relerr(out[k], ref[k]))


def test_rrs_workspace_limit_chunks_owner_range_bit_identically():
    """A small workspace limit makes the library run the owner range in sub-ranges: same bits."""
    from vsmom_b200 import _lib
    m, rs = _scene(nS=48)
    full = raman.rt_run_rrs(m, rs)
    ctx = _lib.Context(0)
    ctx.set_workspace_limit(10 << 20)            # elastic records ~6.3 MB + 135 KB of unit state per owner point -> 2 sub-ranges
    part = raman.rt_run_rrs(m, rs, ctx=ctx)
    ctx.close()
    for k in KEYS:
        assert np.array_equal(part[k], full[k]), k


@pytest.mark.parametrize("devices", [[0], [0, 1]])
def test_rrs_multi_context_in_process_equals_single(devices):
    """vsm_rt_column_rrs_sharded_f64: several GPUs from ONE process (the Julia situation), bit-identical to one device."""
    import torch
    from vsmom_b200 import _lib
    if torch.cuda.device_count() < len(devices):
        pytest.skip("needs %d GPUs" % len(devices))
    m, rs = _scene(nS=70, aerosol=True)
    single = raman.rt_run_rrs(m, rs)
    mc = _lib.MultiContext(devices)
    try:
        multi = raman.rt_run_rrs_multi(m, rs, mc)
    finally:
        mc.close()
    for k in KEYS:
        assert np.array_equal(multi[k], single[k]), k


def test_rrs_no_donor_in_band_gives_elastic_only():
    """Every Raman shift points outside a short band: no unit is valid, the inelastic outputs are exactly zero and the
    elastic ones match the oracle (get_n0_n1 returns empty ranges, inelastic_helper.jl:19-26)."""
    m, rs = _scene(nS=6)
    rs = raman.RRS(np.array([-9, 7, 11], np.int32), rs.varpi_raman[:3], rs.greek_raman)
    ref = _oracle(m, rs)
    out = raman.rt_run_rrs(m, rs)
    assert np.all(out["ieR_SFI"] == 0) and np.all(out["ieT_SFI"] == 0)
    assert np.all(ref["ieR_SFI"] == 0)
    for k in ("R_SFI", "T_SFI"):
        assert relerr(out[k], ref[k]) <= TOL


@pytest.mark.parametrize("aerosol", [False, True])
def test_rt_run_ss_rrs_matches_oracle(aerosol):
    """Single scattering with rotational Raman (rt_run_ss(::RRS)): all six outputs vs the oracle; SS inelastic <= full inelastic."""
    import vsmom_oracle as orc
    import vsmom_oracle_rrs as orr
    m, rs = _scene(aerosol=aerosol, seed=13)
    g = rs.greek_raman
    og = orc.GreekCoefs(*(np.asarray(x, float) for x in (g.alpha, g.beta, g.gamma, g.delta, g.epsilon, g.zeta)))
    ref = orr.rt_run_ss_rrs(to_oracle(m), orr.RRS(np.asarray(rs.i_shift), np.asarray(rs.varpi_raman), og))
    out = raman.rt_run_ss_rrs(m, rs)
    for k in ("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI", "hem_R", "hem_T"):
        assert relerr(out[k], ref[k]) <= TOL, (k, relerr(out[k], ref[k]))
    full = raman.rt_run_rrs(m, rs)
    assert np.all(out["ieR_SFI"][:, 0] <= full["ieR_SFI"][:, 0] * (1 + 1e-9) + 1e-14)

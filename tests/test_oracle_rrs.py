"""Pins the rotational-Raman oracle (oracle/vsmom_oracle_rrs.py).  The reference holds no readable numeric table
for this path (test/test_forward_raman.jl asserts invariants only), so the pin is the closure of the first-order
inelastic terms against the golden-pinned elastic oracle, plus the reference's own invariants."""
import numpy as np
import pytest

import vsmom_oracle as orc
import vsmom_oracle_rrs as rrs


def _greek_mix(a, b, eps):
    f = lambda x, y: (1 - eps) * np.asarray(x, float) + eps * np.asarray(y, float)
    return orc.GreekCoefs(f(a.alpha, b.alpha), f(a.beta, b.beta), f(a.gamma, b.gamma), f(a.delta, b.delta),
                          f(a.epsilon, b.epsilon), f(a.zeta, b.zeta))


def _flat_model(eps, pol_n=3, nS=12, Nz=3, aerosol=False, greek_cab=None, albedo=0.2, tau=0.12):
    q = orc.rt_set_streams_gauss(3, 35.0, [20.0, 50.0], pol_n)
    aer = []
    tau_aer = np.zeros((0, Nz))
    if aerosol:
        g = 0.6
        l = np.arange(6)
        b = (2 * l + 1) * g ** l
        aer = [orc.Aerosol(orc.GreekCoefs(b.copy(), b.copy(), np.zeros(6), b.copy(), np.zeros(6), b.copy()), 0.9, 0.0)]
        tau_aer = np.array([[0.0, 0.05, 0.1][:Nz]])
    return orc.Model(pol_n=pol_n, quad=q, vza=[20.0, 50.0], vaz=[30.0, 140.0],
                     tau_rayl=np.full((nS, Nz), tau) * np.linspace(0.6, 1.4, Nz)[None, :],
                     tau_abs=np.zeros((nS, Nz)), tau_aer=tau_aer, aerosols=aer,
                     greek_rayleigh=greek_cab if greek_cab is not None else orc.get_greek_rayleigh(0.007),
                     albedo=albedo, m_max=5 if aerosol else 2, varpi_cabannes=1.0 - eps)


SHIFTS = np.array([-3, -1, 2, 4])
WEIGHTS = np.array([0.2, 0.4, 0.3, 0.1])


def _closure_error(eps, **kw):
    m = _flat_model(eps, **kw)
    gr = rrs.get_greek_raman(0.75 / 1.0 * 0.1)     # any depolarisation: only the closure matters
    r = rrs.RRS(SHIFTS, rrs.normalize_raman_weights(WEIGHTS, 1.0 - eps), gr)
    out = rrs.rt_run_rrs(m, r)
    tot = _flat_model(eps, greek_cab=_greek_mix(m.greek_rayleigh, gr, eps), **kw)
    tot.varpi_cabannes = 1.0
    ref = orc.rt_run(tot)
    interior = slice(4, m.nSpec - 4)                 # every n0 = n1 + shift inside the band
    errs = []
    for k_el, k_ie in (("R_SFI", "ieR_SFI"), ("T_SFI", "ieT_SFI")):
        a = out[k_el][:, :, interior] + out[k_ie][:, :, interior]
        b = ref[k_el][:, :, interior]
        errs.append(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    return max(errs), out, ref


@pytest.mark.parametrize("aerosol", [False, True])
def test_first_order_closure_flat_spectrum(aerosol):
    e1, out1, _ = _closure_error(1e-2, aerosol=aerosol)
    e2, _, _ = _closure_error(1e-3, aerosol=aerosol)
    assert np.all(np.isfinite(out1["ieR_SFI"]))
    # second-order remainder: quadratic in the Raman fraction
    assert e1 < 2e-4, e1
    assert e2 < 2e-6, e2
    assert 50 < e1 / e2 < 200, (e1, e2)
    # and the inelastic part itself is far larger than the remainder (the closure is not vacuous)
    ie = np.max(np.abs(out1["ieR_SFI"])) / np.max(np.abs(out1["R_SFI"]))
    assert ie > 20 * e1


def test_scalar_closure():
    e1, _, _ = _closure_error(1e-2, pol_n=1)
    e2, _, _ = _closure_error(1e-3, pol_n=1)
    assert e1 < 2e-4 and 50 < e1 / e2 < 200


def test_index_coupling_with_structured_solar_spectrum():
    """Flat atmosphere, non-flat F0: everything but F0[n0] is spectrally constant, so
    ie[n1] = sum_dn (w_dn / sum w) F0[n1 + shift_dn] * ie_flat, including the band-edge truncation of get_n0_n1."""
    eps = 2e-2
    m = _flat_model(eps, albedo=0.0)      # (the Lambertian surface source of the reference carries I0, not F0)
    gr = rrs.get_greek_raman(0.1)
    w = rrs.normalize_raman_weights(WEIGHTS, 1.0 - eps)
    flat = rrs.rt_run_rrs(m, rrs.RRS(SHIFTS, w, gr))
    nS = m.nSpec
    rng = np.random.default_rng(3)
    F0 = np.zeros((3, nS))
    F0[0] = 1.0 + 0.5 * rng.random(nS)
    m2 = _flat_model(eps, albedo=0.0)
    m2.F0 = F0
    out = rrs.rt_run_rrs(m2, rrs.RRS(SHIFTS, w, gr))
    # elastic part scales with F0[n]
    np.testing.assert_allclose(out["R_SFI"], flat["R_SFI"] * F0[0][None, None, :], rtol=1e-12, atol=1e-15)
    unit = flat["ieR_SFI"][:, :, 5] / np.sum(w)          # interior point: all four shifts in range
    expect = np.zeros_like(out["ieR_SFI"])
    for n1 in range(nS):
        for dn, s in enumerate(SHIFTS):
            if 0 <= n1 + s < nS:
                expect[:, :, n1] += w[dn] * F0[0, n1 + s] * unit
    np.testing.assert_allclose(out["ieR_SFI"], expect, rtol=1e-10, atol=1e-14)


def test_reference_invariants_structured_band():
    """test/test_forward_raman.jl:90-188: finite, elastic I > 0, total I > 0, |Q|,|U| <= I, bounded Ring effect."""
    rng = np.random.default_rng(11)
    nS, Nz = 40, 4
    eps = 0.03
    m = _flat_model(eps, nS=nS, Nz=Nz, tau=0.03)
    lines = np.zeros(nS)
    for c in rng.integers(0, nS, 6):
        lines += 3.0 * np.exp(-0.5 * ((np.arange(nS) - c) / 1.2) ** 2)
    m.tau_abs = lines[:, None] * np.linspace(0.05, 0.4, Nz)[None, :]
    shifts = np.array([-9, -6, -4, -2, 2, 4, 6, 9])
    w = rrs.normalize_raman_weights(rng.random(8) + 0.2, 1.0 - eps)
    out = rrs.rt_run_rrs(m, rrs.RRS(shifts, w, rrs.get_greek_raman(0.1)))
    nors = _flat_model(0.0, nS=nS, Nz=Nz, tau=0.03, greek_cab=orc.get_greek_rayleigh(0.028))
    nors.tau_abs = m.tau_abs
    ref = orc.rt_run(nors)
    for k in out:
        assert np.all(np.isfinite(out[k]))
    I_tot = out["R_SFI"][:, 0] + out["ieR_SFI"][:, 0]
    assert np.all(out["R_SFI"][:, 0] > 0) and np.all(I_tot > 0)
    for s in (1, 2):
        assert np.all(np.abs(out["R_SFI"][:, s] + out["ieR_SFI"][:, s]) <= I_tot + 1e-10)
    ring = 100.0 * (I_tot - ref["R_SFI"][:, 0]) / ref["R_SFI"][:, 0]
    assert np.mean(np.abs(ring)) < 50.0 and np.max(np.abs(ring)) < 200.0
    # filling-in: the inelastic share is larger inside the absorption lines than in the continuum
    share = out["ieR_SFI"][0, 0] / I_tot[0]
    inner = slice(9, nS - 9)
    deep = np.argmax(lines[inner]) + 9
    cont = np.argmin(lines[inner]) + 9
    assert share[deep] > share[cont]


def test_get_n0_n1_matches_reference_ranges():
    # inelastic_helper.jl:19-26 (1-based inclusive there, 0-based half-open here)
    for nS, d in ((10, 3), (10, -4), (10, 0), (5, 7), (5, -5)):
        s0, s1 = rrs.get_n0_n1(nS, d)
        lo1, hi1 = max(1, 1 - d), min(nS, nS - d)
        if hi1 < lo1:
            assert s1.stop - s1.start == 0
            continue
        assert (s1.start + 1, s1.stop) == (lo1, hi1)
        assert (s0.start + 1, s0.stop) == (lo1 + d, hi1 + d)


def test_pointwise_evaluation_equals_full_band():
    """rt_run_rrs_points (unit-wise two-point bands) reproduces the full-band oracle at the selected indices."""
    rng = np.random.default_rng(2)
    nS, Nz, eps = 30, 3, 0.03
    m = _flat_model(eps, nS=nS, Nz=Nz, tau=0.05)
    m.tau_abs = rng.random((nS, Nz)) * 0.2
    m.tau_rayl = m.tau_rayl * (1 + 0.1 * rng.random((nS, 1)))
    shifts = np.array([-7, -3, 2, 5, 12])
    w = rrs.normalize_raman_weights(rng.random(5) + 0.2, 1.0 - eps)
    r = rrs.RRS(shifts, w, rrs.get_greek_raman(0.1))
    full = rrs.rt_run_rrs(m, r)
    idx = np.array([0, 4, 13, 22, 29])
    part = rrs.rt_run_rrs_points(m, r, idx)
    for k in full:
        np.testing.assert_allclose(part[k], full[k][:, :, idx], rtol=1e-11, atol=1e-16)


def test_single_scatter_rrs_closure_and_thin_limit():
    """rt_run_ss(::RRS): (i) in the flat-spectrum case elastic + inelastic single scattering equals the elastic single
    scattering of the mixed phase matrix EXACTLY (single scattering is linear in the phase matrix: no O(eps^2) remainder);
    (ii) for a thin atmosphere the inelastic part converges to the full RRS solution (gap linear in the optical depth)."""
    eps = 2e-2
    m = _flat_model(eps, albedo=0.0, tau=0.02)
    gr = rrs.get_greek_raman(0.1)
    r = rrs.RRS(SHIFTS, rrs.normalize_raman_weights(WEIGHTS, 1.0 - eps), gr)
    ss = rrs.rt_run_ss_rrs(m, r)
    tot = _flat_model(eps, albedo=0.0, tau=0.02, greek_cab=_greek_mix(m.greek_rayleigh, gr, eps))
    tot.varpi_cabannes = 1.0
    ref = orc.rt_run_ss(tot)
    inner = slice(4, m.nSpec - 4)
    for a, b in (("R_SFI", "ieR_SFI"), ("T_SFI", "ieT_SFI")):
        np.testing.assert_allclose(ss[a][:, :, inner] + ss[b][:, :, inner], ref[a][:, :, inner], rtol=1e-11, atol=1e-16)
    gaps = []
    for tau in (4e-3, 2e-3):
        mm = _flat_model(eps, albedo=0.0, tau=tau)
        full = rrs.rt_run_rrs(mm, r)
        s1 = rrs.rt_run_ss_rrs(mm, r)
        gaps.append(np.max(np.abs(full["ieR_SFI"][:, 0, inner] - s1["ieR_SFI"][:, 0, inner]) / np.abs(full["ieR_SFI"][:, 0, inner])))
    assert gaps[0] < 5e-2 and 1.6 < gaps[0] / gaps[1] < 2.4, gaps


def test_outputs_do_not_read_composite_ieRmp_ieTpp(monkeypatch):
    """DESIGN.md §3.8, inelastic half (interaction_inelastic.jl:372-502): ieR-+ reads ieT++, the elastic T++ and itself, ieT++
    reads itself, and no ieJ0+- / ieT-- / ieR+- line reads either -- nor the elastic R-+ / T++.  Overwriting all four with noise
    after every interaction leaves the elastic and inelastic outputs unchanged bit for bit."""
    eps = 0.03
    m = _flat_model(eps, nS=16, Nz=3, tau=0.05)
    m.tau_abs = np.linspace(0.0, 0.3, 16)[:, None] * np.ones(3)[None, :]
    shifts = np.array([-5, -2, 3, 6])
    w = rrs.normalize_raman_weights(np.array([0.3, 0.2, 0.4, 0.1]), 1.0 - eps)
    R = rrs.RRS(shifts, w, rrs.get_greek_raman(0.1))
    ref = rrs.rt_run_rrs(m, R)
    plain = rrs.interaction_inelastic_11
    rng = np.random.default_rng(5)
    calls = []

    def poisoned(*args, **kw):
        C, CI = plain(*args, **kw)
        C.Rmp = rng.standard_normal(C.Rmp.shape)
        C.Tpp = rng.standard_normal(C.Tpp.shape)
        CI.ieRmp = rng.standard_normal(CI.ieRmp.shape)
        CI.ieTpp = rng.standard_normal(CI.ieTpp.shape)
        calls.append(1)
        return C, CI

    monkeypatch.setattr(rrs, "interaction_inelastic_11", poisoned)
    out = rrs.rt_run_rrs(m, R)
    assert len(calls) >= 3
    for k in ("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI"):
        assert np.array_equal(out[k], ref[k]), k

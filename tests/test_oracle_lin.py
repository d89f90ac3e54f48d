"""Oracle-only (CPU) tests of the tangent-linear doubling / interaction restatement
(doubling_lin.jl:216-335, interaction_lin.jl:54-331) against central finite differences of
the forward oracle -- the reference pins its own LinMode the same way
(test/test_forward_lin.jl:167-244, analytic vs FD)."""
import copy

import numpy as np
import pytest

import vsmom_oracle as orc
import vsmom_b200 as vb
from helpers import to_oracle


def _layer(nS=3, seed=0):
    """A physically plausible homogeneous layer (elemental r, t, j) from the forward oracle."""
    m = vb.synthetic_o2a_model(nSpec=nS, nstreams=3, polarization_type="Stokes_IQU", aerosol=True)
    om = to_oracle(m)
    ctl = orc.band_control(om, 0)
    props = orc.construct_core_optical_properties(om, 0)
    iz = int(np.argmax(ctl.ndoubl))
    nd = int(ctl.ndoubl[iz])
    p = props[iz]
    dtau = p.tau / 2.0 ** nd
    F0 = np.zeros((om.pol_n, nS)); F0[0] = 1
    r, t, _, _, j0p, j0m = orc.elemental(om, 0, nd, dtau, p.varpi, p.Zpp, p.Zmp, ctl.tau_sum[:, iz], F0)
    return om, nd, dtau, r, t, j0p, j0m


def _rand_like(rng, x, P, scale):
    return scale * rng.standard_normal((P,) + x.shape) * (np.abs(x).max() + 1e-30)


def test_doubling_lin_matches_finite_differences():
    om, nd, dtau, r, t, j0p, j0m = _layer()
    nd = min(nd, 6)
    rng = np.random.default_rng(5)
    P = 2
    dr, dt = _rand_like(rng, r, P, 0.3), _rand_like(rng, t, P, 1e-3)
    djp, djm = _rand_like(rng, j0p, P, 0.3), _rand_like(rng, j0m, P, 0.3)
    dtd = 0.2 * rng.standard_normal((P, dtau.size)) * dtau[None]
    mu0 = om.quad.mu0
    fwd, lin = orc.doubling_lin(om, nd, np.exp(-dtau / mu0), r, t, j0p, j0m, dr, dt, djp, djm, dtd)
    ref = orc.doubling(om, nd, np.exp(-dtau / mu0), r, t, j0p, j0m)
    for a, b in zip(fwd, ref):
        assert np.max(np.abs(a - b)) <= 1e-14 * max(np.abs(b).max(), 1e-300)
    h = 1e-6
    for ip in range(P):
        out = []
        for sg in (+1, -1):
            e = np.exp(-(dtau + sg * h * dtd[ip]) / mu0)
            out.append(orc.doubling(om, nd, e, r + sg * h * dr[ip], t + sg * h * dt[ip], j0p + sg * h * djp[ip],
                                    j0m + sg * h * djm[ip]))
        for k, (up, dn) in enumerate(zip(*out)):
            fd = (up - dn) / (2 * h)
            scale = max(np.abs(fd).max(), 1e-300)
            assert np.max(np.abs(lin[k][ip] - fd)) <= 2e-7 * scale, (ip, k)


@pytest.mark.parametrize("itf", [0, 1, 2, 3])
def test_interaction_lin_matches_finite_differences(itf):
    om, nd, dtau, r, t, j0p, j0m = _layer(seed=1)
    mu0 = om.quad.mu0
    e = np.exp(-dtau / mu0)
    a = orc.doubling(om, min(nd, 5), e, r, t, j0p, j0m)            # added layer (r-+, t++, r+-, t--, j+, j-)
    A = (a[0], a[2], a[1], a[3], a[4], a[5])                       # interaction() order: r-+, r+-, t++, t--, j+, j-
    B = orc.doubling(om, min(nd, 4), e, 0.7 * r, t, 0.5 * j0p, 0.8 * j0m)   # stands in for the composite
    rng = np.random.default_rng(11 + itf)
    P = 2
    dA = [_rand_like(rng, x, P, 0.2 if k not in (2, 3) else 1e-3) for k, x in enumerate(A)]
    C0 = orc.Composite(B[0].copy(), B[2].copy(), B[1].copy(), B[3].copy(), B[4].copy(), B[5].copy())
    dC0 = orc.Composite(*[_rand_like(rng, x, P, 0.2) for x in (C0.Rmp, C0.Rpm, C0.Tpp, C0.Tmm, C0.J0p, C0.J0m)])
    C1, dC1 = orc.interaction_lin(itf, copy.deepcopy(C0), copy.deepcopy(dC0), *A, *dA)
    Cf = orc.interaction(itf, copy.deepcopy(C0), *A)
    fields = ("Rmp", "Rpm", "Tpp", "Tmm", "J0p", "J0m")
    for f in fields:
        assert np.max(np.abs(getattr(C1, f) - getattr(Cf, f))) <= 1e-14 * max(np.abs(getattr(Cf, f)).max(), 1e-300), f
    h = 1e-6
    for ip in range(P):
        res = []
        for sg in (+1, -1):
            Cp = orc.Composite(*[getattr(C0, f) + sg * h * getattr(dC0, f)[ip] for f in fields])
            Ap = [x + sg * h * dx[ip] for x, dx in zip(A, dA)]
            res.append(orc.interaction(itf, Cp, *Ap))
        for f in fields:
            fd = (getattr(res[0], f) - getattr(res[1], f)) / (2 * h)
            scale = max(np.abs(fd).max(), 1e-300)
            assert np.max(np.abs(getattr(dC1, f)[ip] - fd)) <= 2e-7 * scale, (itf, ip, f)


@pytest.mark.parametrize("nd_mode", ["doubled", "nd0"])
def test_elemental_lin_matches_finite_differences(nd_mode):
    m = vb.synthetic_o2a_model(nSpec=3, nstreams=3, polarization_type="Stokes_IQU", aerosol=True)
    om = to_oracle(m)
    ctl = orc.band_control(om, 1)
    props = orc.construct_core_optical_properties(om, 1)
    iz = int(np.argmax(ctl.ndoubl))
    nd = int(ctl.ndoubl[iz]) if nd_mode == "doubled" else 0
    p = props[iz]
    nS = 3
    dtau = p.tau / 2.0 ** max(nd, 4)
    F0 = np.zeros((om.pol_n, nS)); F0[0] = 1; F0[1] = 0.1
    tsum = ctl.tau_sum[:, iz] + 0.05
    rng = np.random.default_rng(3)
    P = 2
    N = om.quad.qp_muN.size
    Zpp = np.broadcast_to(p.Zpp, (nS, N, N)).copy(); Zmp = np.broadcast_to(p.Zmp, (nS, N, N)).copy()
    dd = 0.3 * rng.standard_normal((P, nS)) * dtau[None]
    dw = 0.3 * rng.standard_normal((P, nS)) * p.varpi[None]
    dZp = 0.3 * rng.standard_normal((P, nS, N, N)) * np.abs(Zpp).max()
    dZm = 0.3 * rng.standard_normal((P, nS, N, N)) * np.abs(Zmp).max()
    dts = 0.3 * rng.standard_normal((P, nS)) * tsum[None]
    lin = orc.elemental_lin(om, 1, nd, dtau, p.varpi, Zpp, Zmp, tsum, F0, dd, dw, dZp, dZm, dts)
    h = 1e-6
    for ip in range(P):
        res = [orc.elemental(om, 1, nd, dtau + sg * h * dd[ip], p.varpi + sg * h * dw[ip], Zpp + sg * h * dZp[ip],
                             Zmp + sg * h * dZm[ip], tsum + sg * h * dts[ip], F0) for sg in (+1, -1)]
        for k, (up, dn) in enumerate(zip(*res)):
            if up is None:
                assert lin[k] is None
                continue
            fd = (up - dn) / (2 * h)
            scale = max(np.abs(fd).max(), 1e-300)
            assert np.max(np.abs(lin[k][ip] - fd)) <= 5e-7 * scale, (nd_mode, ip, k)


def test_rt_run_lin_jacobian_matches_finite_differences_of_rt_run():
    """Column-level Jacobian (gas column, aerosol optical depth, Rayleigh scaling, Lambertian albedo)
    against central differences of the forward oracle; tolerance far inside the reference's own
    analytic-vs-FD bounds (1 % - 20 %, test/test_forward_lin.jl:167-244)."""
    m = vb.synthetic_o2a_model(nSpec=4, nstreams=3, polarization_type="Stokes_IQU", aerosol=True, albedo=0.2)
    m.tau_abs = m.tau_abs * np.logspace(-4.0, -1.5, 4)[:, None]
    om = to_oracle(m)
    nS, Nz, nAer = om.nSpec, om.Nz, len(om.aerosols)
    P = 4
    z3 = np.zeros((P, nS, Nz)); za = np.zeros((P, nAer, Nz))
    d_rayl, d_abs, d_aer, d_alb = z3.copy(), z3.copy(), za.copy(), np.zeros(P)
    d_abs[0] = om.tau_abs            # p0: gas column scaling
    d_aer[1] = om.tau_aer            # p1: aerosol optical depth scaling
    d_rayl[2] = om.tau_rayl          # p2: Rayleigh scaling
    d_alb[3] = 1.0                   # p3: albedo
    lin = orc.rt_run_lin(om, d_rayl, d_abs, d_aer, d_alb)
    fwd = orc.rt_run(om)
    for k in ("R_SFI", "T_SFI"):
        assert np.max(np.abs(lin[k] - fwd[k])) <= 1e-13 * np.abs(fwd[k]).max()
    h = 1e-5
    for ip in range(P):
        res = []
        for sg in (+1, -1):
            mp = copy.deepcopy(om)
            mp.tau_abs = om.tau_abs + sg * h * d_abs[ip]
            mp.tau_aer = om.tau_aer + sg * h * d_aer[ip]
            mp.tau_rayl = om.tau_rayl + sg * h * d_rayl[ip]
            mp.albedo = om.albedo + sg * h * d_alb[ip]
            res.append(orc.rt_run(mp))
        for k in ("R_SFI", "T_SFI"):
            fd = (res[0][k] - res[1][k]) / (2 * h)
            scale = max(np.abs(fd).max(), 1e-300)
            err = np.max(np.abs(lin["d" + k][ip] - fd)) / scale
            assert err <= 1e-6, (ip, k, err)


def test_lin_outputs_do_not_read_composite_Rmp_Tpp(monkeypatch):
    """DESIGN.md §3.8, LinMode (interaction_lin.jl:238-306): the composite R-+ / T++ and their derivatives feed only
    themselves; overwriting them with noise after every interaction leaves the forward outputs and the Jacobians unchanged
    bit for bit (k_adding_lin does not advance them)."""
    m = vb.synthetic_o2a_model(nSpec=3, nstreams=3, polarization_type="Stokes_IQU", aerosol=True, albedo=0.2)
    om = to_oracle(m)
    nS, Nz, nAer = om.nSpec, om.Nz, len(om.aerosols)
    P = 2
    d_rayl = np.zeros((P, nS, Nz)); d_abs = np.zeros((P, nS, Nz)); d_aer = np.zeros((P, nAer, Nz)); d_alb = np.array([0.0, 1.0])
    d_abs[0] = om.tau_abs
    ref = orc.rt_run_lin(om, d_rayl, d_abs, d_aer, d_alb)
    plain = orc.interaction_lin
    rng = np.random.default_rng(9)
    calls = []

    def poisoned(*args, **kw):
        C, dC = plain(*args, **kw)
        C.Rmp = rng.standard_normal(C.Rmp.shape); C.Tpp = rng.standard_normal(C.Tpp.shape)
        dC.Rmp = rng.standard_normal(dC.Rmp.shape); dC.Tpp = rng.standard_normal(dC.Tpp.shape)
        calls.append(1)
        return C, dC

    monkeypatch.setattr(orc, "interaction_lin", poisoned)
    out = orc.rt_run_lin(om, d_rayl, d_abs, d_aer, d_alb)
    assert len(calls) > 2
    for k in ("R_SFI", "T_SFI", "dR_SFI", "dT_SFI"):
        assert np.array_equal(out[k], ref[k]), k

"""
CPU oracle for the rotational-Raman (RRS) variant of the CoreRT hot path (BASELINE config 5).

TEST INFRASTRUCTURE ONLY -- same rule as vsmom_oracle.py: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

NumPy restatement (batch-first arrays: ``(nRaman, nSpec, N, N)`` for the reference's
``(N, N, nSpec, nRaman)``) of

* rt_kernel!(::RRS, ...)             src/CoreRT/CoreKernel/rt_kernel.jl:352-391
* elemental_inelastic!  (+ kernels)  src/CoreRT/CoreKernel/elemental_inelastic.jl:23-206, 479-603, 613-632
* doubling_helper!(::RRS)            src/CoreRT/CoreKernel/doubling_inelastic.jl:13-164, 335-358, 396-408
* interaction_helper!(::RRS, ::ScatteringInterface_11)   src/CoreRT/CoreKernel/interaction_inelastic.jl:319-521
* copy_added_to_composite_ie!        src/CoreRT/CoreKernel/rt_helpers.jl:222-228
* postprocessing_vza!(::RRS, ...)    src/CoreRT/tools/postprocessing_vza.jl:121-162
* the RRS driver loop                src/CoreRT/rt_run.jl:238-539, constructCoreOpticalProperties
                                     (compEffectiveLayerProperties.jl:11-65: fScattRayleigh)
* get_n₀_n₁                          src/Inelastic/inelastic_helper.jl:19-26
* get_greek_raman                    src/Inelastic/inelastic_helper.jl:862-880
* normalize_raman_weights!           src/Inelastic/types.jl:625-629

The Raman *physics* producers (N2/O2 molecular constants, line strengths -> ϖ_λ₁λ₀ and i_λ₁λ₀,
getRamanSSProp!, raman_atmo_prop.jl:80-100) are out of scope: the shift table and weights are inputs.

PARITY PINNING.  The reference pins this path only with invariants (test/test_forward_raman.jl: finite,
positive, |Q|,|U| <= I, bounded Ring effect) and a JLD2 file that cannot be read in this container, so the
numeric pin used here (tests/test_oracle_rrs.py) is a closure property of the algorithm itself: the
inelastic terms are the first-order (single Raman event) perturbation, so for a spectrally flat atmosphere
and solar spectrum   elastic(ϖ_Cabannes = 1-ε, Z_Cabannes) + inelastic   must equal the golden-pinned elastic
oracle run with the phase matrix (1-ε) Z_Cabannes + ε Z_Raman up to O(ε²) -- checked at two ε with the
error ratio ≈ 100 -- plus the reference's invariants and the index-coupling check with a non-flat F₀.
Every layer must scatter (the reference's RRS rt_kernel! hard-codes ``scatter = true``,
rt_kernel.jl:366; its _01/_10 RRS interaction helpers reference an undefined ``ieJ₁⁺``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

import vsmom_oracle as orc
from vsmom_oracle import GreekCoefs, Model, Composite, _D_diag, _mv, cosd, sind, nearest_point, expdiff_neg

RT_WEIGHT_TOL = 1e-8     # rt_helpers.jl:58
RT_CLOSE_TOL = 1e-8      # rt_helpers.jl:68
RT_LOOSE_TOL = 1e-6      # rt_helpers.jl:76-78


@dataclass
class RRS:
    """The fields of InelasticScattering.RRS the hot path reads (src/Inelastic/types.jl)."""
    i_shift: np.ndarray        # i_λ₁λ₀ [nRaman] (ints): n₀ = n₁ + i_shift[Δn]
    varpi_raman: np.ndarray    # ϖ_λ₁λ₀ [nRaman], already normalised (normalize_raman_weights!)
    greek_raman: GreekCoefs


def get_greek_raman(depol: float) -> GreekCoefs:
    """inelastic_helper.jl:862-880 (depol = rho_depol_RotRaman)."""
    dpl_p = (1 - depol) / (1 + depol / 2)
    dpl_r = (1 - 2 * depol) / (1 - depol)
    return GreekCoefs(alpha=np.array([0.0, 0.0, 3 * dpl_p]), beta=np.array([1.0, 0.0, 0.5 * dpl_p]),
                      gamma=np.array([0.0, 0.0, dpl_p * math.sqrt(1.5)]), delta=np.array([0.0, dpl_p * dpl_r * 1.5, 0.0]),
                      epsilon=np.zeros(3), zeta=np.zeros(3))


def normalize_raman_weights(varpi_raman, varpi_cabannes):
    """types.jl:625-629"""
    w = np.asarray(varpi_raman, float)
    return w * ((1 - varpi_cabannes) / np.sum(w))


def get_n0_n1(nSpec: int, shift: int):
    """inelastic_helper.jl:19-26 (0-based half-open python ranges)."""
    lo = max(0, -shift)
    hi = min(nSpec, nSpec - shift)
    if hi <= lo:
        return slice(0, 0), slice(0, 0)
    return slice(lo + shift, hi + shift), slice(lo, hi)


def fscatt_rayleigh(model: Model):
    """fScattRayleigh = rayl.τ ./ combo.τ before the absorption is added (compEffectiveLayerProperties.jl:56)."""
    FT = model.FT
    out = np.empty((model.nSpec, model.Nz), FT)
    for iz in range(model.Nz):
        tr = model.tau_rayl[:, iz].astype(FT)
        tot = tr.copy()
        for ia, a in enumerate(model.aerosols):
            tot = tot + FT((1 - a.ft * a.omega) * model.tau_aer[ia, iz])
        out[:, iz] = tr / tot
    return out


# --------------------------------------------------------------------------
# elemental_inelastic!   elemental_inelastic.jl:23-206 (RT), :479-603 (SFI), :613-632 (D)
# --------------------------------------------------------------------------
def elemental_inelastic(model: Model, rrs: RRS, m: int, ndoubl: int, dtau, fscatt, tau_sum, F0, Zpp, Zmp):
    """dtau, fscatt, tau_sum: (nS,) of the FULL band;  F0 (pol_n, nS);  Zpp/Zmp: the Raman Z moments (N,N).
    Returns ier_mp, iet_pp, ier_pm, iet_mm (the last two only valid for ndoubl == 0), ieJ0p, ieJ0m."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    mu = q.qp_muN.astype(FT)
    N = mu.size
    nS = dtau.shape[0]
    nR = len(rrs.i_shift)
    wct02 = FT(0.5 if m == 0 else 0.25)
    wct2 = (q.wt_muN / 2 if m == 0 else q.wt_muN / 4).astype(FT)
    ier = np.zeros((nR, nS, N, N), FT)
    iet = np.zeros((nR, nS, N, N), FT)
    ieJp = np.zeros((nR, nS, N), FT)
    ieJm = np.zeros((nR, nS, N), FT)
    mui = mu[None, :, None]
    muj = mu[None, None, :]
    same = (mui == muj)
    active = (wct2 > RT_WEIGHT_TOL)[None, None, :]
    i_start = pol_n * q.imu0
    i_end = i_start + pol_n
    mu_s = mu[i_start]
    insun = np.zeros(N, bool)
    insun[i_start:i_end] = True
    Dd = _D_diag(pol_n, q.Nquad, FT)
    for dn in range(nR):
        s0, s1 = get_n0_n1(nS, int(rrs.i_shift[dn]))
        if s1.stop <= s1.start:
            continue
        d1 = dtau[s1][:, None, None]
        d0 = dtau[s0][:, None, None]
        fw = (fscatt[s0] * FT(rrs.varpi_raman[dn]))[:, None, None]
        ratio = d1 / d0
        # ier⁻⁺ (:139-143)
        r = fw * Zmp[None] * (1 / ((mui / muj) + ratio)) * (-np.expm1(-((d1 / mui) + (d0 / muj)))) * wct2[None, None, :]
        # iet⁺⁺ (:146-196)
        with np.errstate(divide='ignore', invalid='ignore'):
            ed = expdiff_neg(d1 / mui, d0 / muj)
            t_same_far = fw * Zpp[None] * wct2[None, None, :] * ed / (1 - ratio)
            t_lin = (d0 / mui) * fw * Zpp[None] * wct2[None, None, :] * np.exp(-d0 / muj)
            t_off = fw * Zpp[None] * (1 / ((mui / muj) - ratio)) * wct2[None, None, :] * ed
        far = np.abs(d0 - d1) > RT_LOOSE_TOL
        close = np.abs((mui / muj) - ratio) < RT_CLOSE_TOL
        t = np.where(same, np.where(far, t_same_far, t_lin), np.where(close, t_lin, t_off))
        ier[dn, s1] = np.where(active, r, 0)
        iet[dn, s1] = np.where(active, t, 0)
        # sources (:513-589)
        F0n = F0[:, s0].astype(FT)                                    # (pol_n, n)
        ZppI = np.einsum('ik,kn->ni', Zpp[:, i_start:i_end], F0n)
        ZmpI = np.einsum('ik,kn->ni', Zmp[:, i_start:i_end], F0n)
        dd1 = dtau[s1][:, None]
        dd0 = dtau[s0][:, None]
        fw1 = (fscatt[s0] * FT(rrs.varpi_raman[dn]))[:, None]
        rat = dd1 / dd0
        mu_i = mu[None, :]
        with np.errstate(divide='ignore', invalid='ignore'):
            jp_in_far = fw1 * ZppI * wct02 * expdiff_neg(dd1 / mu_i, dd0 / mu_i) / (1 - rat)
            jp_in_lin = (dd0 / mu_i) * wct02 * fw1 * ZppI * np.exp(-dd0 / mu_i)
            jp_out_lin = (dd0 / mu_i) * wct02 * fw1 * ZppI * np.exp(-dd0 / mu_s)
            jp_out = wct02 * fw1 * ZppI * (1 / ((mu_i / mu_s) - rat)) * expdiff_neg(dd1 / mu_i, dd0 / mu_s)
        far1 = np.abs(dd0 - dd1) > RT_CLOSE_TOL
        close1 = np.abs((mu_i / mu_s) - rat) < RT_CLOSE_TOL
        jp = np.where(insun[None, :], np.where(far1, jp_in_far, jp_in_lin), np.where(close1, jp_out_lin, jp_out))
        jm = wct02 * fw1 * ZmpI * (1 / ((mu_i / mu_s) + rat)) * (-np.expm1(-((dd1 / mu_i) + (dd0 / mu_s))))
        att = np.exp(-tau_sum[s0] / mu_s)[:, None]
        jp = jp * att
        jm = jm * att
        if ndoubl >= 1:
            jm = Dd[None, :] * jm
        ieJp[dn, s1] = jp
        ieJm[dn, s1] = jm
    # apply_D_matrix_elemental! (:613-632, 704-722); apply_D_matrix_elemental_SFI! is a no-op for RRS
    ier_pm = iet_mm = None
    if pol_n == 1:
        ier_pm, iet_mm = ier.copy(), iet.copy()
    elif ndoubl < 1:
        sgn = (Dd[:, None] * Dd[None, :])[None, None]
        ier_pm = sgn * ier
        iet_mm = sgn * iet
    else:
        ier = Dd[None, None, :, None] * ier
    return ier, iet, ier_pm, iet_mm, ieJp, ieJm


# --------------------------------------------------------------------------
# doubling_helper!(::RRS)   doubling_inelastic.jl:13-164
# --------------------------------------------------------------------------
def doubling_inelastic(model: Model, rrs: RRS, ndoubl: int, expk, r, t, j0p, j0m, ier, iet, ieJp, ieJm):
    """Elastic r, t (nS,N,N), j0± (nS,N) and inelastic ier, iet (nR,nS,N,N), ieJ± (nR,nS,N) after ndoubl doublings
    and the final apply_D calls.  ndoubl >= 1."""
    FT = model.FT
    N = r.shape[1]
    nS = r.shape[0]
    nR = len(rrs.i_shift)
    I = np.eye(N, dtype=FT)[None]
    expk = expk.copy()
    jp = j0p[:, :, None].copy()
    jm = j0m[:, :, None].copy()
    ieJp = ieJp[..., None].copy()
    ieJm = ieJm[..., None].copy()
    ier = ier.copy()
    iet = iet.copy()
    for _ in range(ndoubl):
        gp = np.linalg.inv(I - r @ r)
        tt = t @ gp
        e = expk[:, None, None]
        j1p = jp * e
        j1m = jm * e
        tmp1 = gp @ (jp + r @ j1m)
        tmp2 = gp @ (j1m + r @ jp)
        for dn in range(nR):
            s0, s1 = get_n0_n1(nS, int(rrs.i_shift[dn]))
            if s1.stop <= s1.start:
                continue
            e0 = expk[s0][:, None, None]
            ie1p = ieJp[dn, s1] * e0
            ie1m = ieJm[dn, s1] * e0
            A = ier[dn, s1]
            B = iet[dn, s1]
            tmp3 = (ie1p + (tt[s1] @ (ieJp[dn, s1] + r[s1] @ ie1m + A @ j1m[s0]
                                      + (r[s1] @ A + A @ r[s0]) @ tmp1[s0])) + B @ tmp1[s0])
            tmp4 = (ieJm[dn, s1] + (tt[s1] @ (ie1m + A @ jp[s0] + r[s1] @ ieJp[dn, s1]
                                              + (A @ r[s0] + r[s1] @ A) @ tmp2[s0])) + B @ tmp2[s0])
            ieJp[dn, s1] = tmp3
            ieJm[dn, s1] = tmp4
        jm_new = jm + tt @ (j1m + r @ jp)
        jp_new = j1p + tt @ (jp + r @ j1m)
        jm, jp = jm_new, jp_new
        expk = expk ** 2
        for dn in range(nR):
            s0, s1 = get_n0_n1(nS, int(rrs.i_shift[dn]))
            if s1.stop <= s1.start:
                continue
            A = ier[dn, s1]
            B = iet[dn, s1]
            X = A @ r[s0] + r[s1] @ A
            tmp5 = tt[s1] @ (B + X @ gp[s0] @ t[s0]) + B @ gp[s0] @ t[s0]
            tmp6 = (A + (B @ gp[s0] @ r[s0] @ t[s0])
                    + tt[s1] @ (r[s1] @ B + (A + X @ gp[s0] @ r[s0]) @ t[s0]))
            iet[dn, s1] = tmp5
            ier[dn, s1] = tmp6
        r_new = r + (tt @ r) @ t
        t = tt @ t
        r = r_new
    Dd = _D_diag(model.pol_n, model.quad.Nquad, FT)
    jp, jm = jp[:, :, 0], jm[:, :, 0]
    ieJp, ieJm = ieJp[..., 0], ieJm[..., 0]
    if model.pol_n == 1:
        return r, t, r.copy(), t.copy(), jp, jm, ier, iet, ier.copy(), iet.copy(), ieJp, ieJm
    sgn = (Dd[:, None] * Dd[None, :])
    r = Dd[None, :, None] * r
    r_pm = sgn[None] * r
    t_mm = sgn[None] * t
    jm = Dd[None, :] * jm
    ier = Dd[None, None, :, None] * ier           # apply_D_IE_RRS! (:335-358)
    ier_pm = sgn[None, None] * ier
    iet_mm = sgn[None, None] * iet
    ieJm = Dd[None, None, :] * ieJm               # apply_D_SFI_IE_RRS! (:396-408)
    return r, t, r_pm, t_mm, jp, jm, ier, iet, ier_pm, iet_mm, ieJp, ieJm


@dataclass
class CompositeIE:
    ieRmp: np.ndarray    # (nR, nS, N, N)
    ieRpm: np.ndarray
    ieTpp: np.ndarray
    ieTmm: np.ndarray
    ieJ0p: np.ndarray    # (nR, nS, N)
    ieJ0m: np.ndarray


# --------------------------------------------------------------------------
# interaction_helper!(::RRS, ::ScatteringInterface_11)   interaction_inelastic.jl:319-521
# --------------------------------------------------------------------------
def interaction_inelastic_11(rrs: RRS, C: Composite, CI: CompositeIE, r_mp, r_pm, t_pp, t_mm, j0p, j0m,
                             ier_mp, ier_pm, iet_pp, iet_mm, a_ieJp, a_ieJm):
    FT = C.Rmp.dtype
    nS, N = C.Rmp.shape[0], C.Rmp.shape[1]
    nR = len(rrs.i_shift)
    I = np.eye(N, dtype=FT)[None]
    v = lambda x: x[..., None]
    J0p, J0m = v(C.J0p), v(C.J0m)
    aj0p, aj0m = v(j0p), v(j0m)
    new = CompositeIE(*(np.zeros_like(x) for x in (CI.ieRmp, CI.ieRpm, CI.ieTpp, CI.ieTmm, CI.ieJ0p, CI.ieJ0m)))
    inv1 = np.linalg.inv(I - r_mp @ C.Rpm)
    T01 = C.Tmm @ inv1
    for dn in range(nR):
        s0, s1 = get_n0_n1(nS, int(rrs.i_shift[dn]))
        if s1.stop <= s1.start:
            continue
        P = T01[s1] @ (ier_mp[dn, s1] @ C.Rpm[s0] + r_mp[s1] @ CI.ieRpm[dn, s1]) + CI.ieTmm[dn, s1]
        new.ieJ0m[dn, s1] = (v(CI.ieJ0m[dn, s1])
                             + T01[s1] @ (ier_mp[dn, s1] @ J0p[s0] + r_mp[s1] @ v(CI.ieJ0p[dn, s1]) + v(a_ieJm[dn, s1]))
                             + P @ inv1[s0] @ (aj0m[s0] + r_mp[s0] @ J0p[s0]))[..., 0]
        new.ieRmp[dn, s1] = (CI.ieRmp[dn, s1]
                             + T01[s1] @ (ier_mp[dn, s1] @ C.Tpp[s0] + r_mp[s1] @ CI.ieTpp[dn, s1])
                             + P @ inv1[s0] @ r_mp[s0] @ C.Tpp[s0])
        new.ieTmm[dn, s1] = T01[s1] @ iet_mm[dn, s1] + P @ inv1[s0] @ t_mm[s0]
    J0m_new = J0m + T01 @ (r_mp @ J0p + aj0m)
    Rmp_new = C.Rmp + T01 @ r_mp @ C.Tpp
    Tmm_new = T01 @ t_mm
    inv2 = np.linalg.inv(I - C.Rpm @ r_mp)
    T21 = t_pp @ inv2
    for dn in range(nR):
        s0, s1 = get_n0_n1(nS, int(rrs.i_shift[dn]))
        if s1.stop <= s1.start:
            continue
        P = T21[s1] @ (CI.ieRpm[dn, s1] @ r_mp[s0] + C.Rpm[s1] @ ier_mp[dn, s1]) + iet_pp[dn, s1]
        new.ieJ0p[dn, s1] = (v(a_ieJp[dn, s1])
                             + T21[s1] @ (v(CI.ieJ0p[dn, s1]) + CI.ieRpm[dn, s1] @ aj0m[s0] + C.Rpm[s1] @ v(a_ieJm[dn, s1]))
                             + P @ inv2[s0] @ (J0p[s0] + C.Rpm[s0] @ aj0m[s0]))[..., 0]
        new.ieTpp[dn, s1] = T21[s1] @ CI.ieTpp[dn, s1] + P @ inv2[s0] @ C.Tpp[s0]
        new.ieRpm[dn, s1] = (ier_pm[dn, s1]
                             + T21[s1] @ (CI.ieRpm[dn, s1] @ t_mm[s0] + C.Rpm[s1] @ iet_mm[dn, s1])
                             + P @ inv2[s0] @ C.Rpm[s0] @ t_mm[s0])
    J0p_new = aj0p + T21 @ (J0p + C.Rpm @ aj0m)
    Tpp_new = T21 @ C.Tpp
    Rpm_new = r_pm + T21 @ C.Rpm @ t_mm
    C.J0m, C.Rmp, C.Tmm = J0m_new[..., 0], Rmp_new, Tmm_new
    C.J0p, C.Tpp, C.Rpm = J0p_new[..., 0], Tpp_new, Rpm_new
    return C, new


# --------------------------------------------------------------------------
# rt_run(::RRS, model, iBand)   rt_run.jl:238-539 with rt_kernel!(::RRS) rt_kernel.jl:352-391
# --------------------------------------------------------------------------
def rt_run_rrs(model: Model, rrs: RRS, return_streams: bool = False, ctl_override=None):
    """Returns dict(R_SFI, T_SFI, ieR_SFI, ieT_SFI), each (nVZA, pol_n, nSpec).  ``model.greek_rayleigh`` must be
    the Cabannes set and ``model.varpi_cabannes`` ϖ_Cabannes (compEffectiveLayerProperties.jl:36-41)."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    nS = model.nSpec
    N = q.qp_muN.size
    nV = len(model.vza)
    nR = len(rrs.i_shift)
    R_SFI = np.zeros((nV, pol_n, nS), FT)
    T_SFI = np.zeros((nV, pol_n, nS), FT)
    ieR_SFI = np.zeros((nV, pol_n, nS), FT)
    ieT_SFI = np.zeros((nV, pol_n, nS), FT)
    F0 = model.F0
    if F0 is None:
        F0 = np.zeros((pol_n, nS), FT)
        F0[0, :] = 1
    fsc = fscatt_rayleigh(model)
    streams = []
    for m in range(model.m_max + 1):
        weight = FT(0.5 / math.pi) if m == 0 else FT(1.0 / math.pi)
        Zpp_r, Zmp_r = orc.compute_Z_moments(pol_n, q.qp_mu, rrs.greek_raman, m, FT)     # computeRamanZλ!
        ctl = orc.band_control(model, m)
        if ctl_override is not None:      # control flow of the FULL band when `model` holds a subset of its points
            ctl = orc.BandControl(ctl_override.scatter, ctl_override.ndoubl, ctl_override.interface, ctl.tau_sum)
        props = orc.construct_core_optical_properties(model, m)
        if not all(ctl.scatter):
            raise ValueError("RRS: every layer must scatter (rt_kernel.jl:366)")
        C = CI = None
        for iz in range(model.Nz):
            p = props[iz]
            nd = ctl.ndoubl[iz]
            dtau = (p.tau / FT(2 ** nd)).astype(FT)
            expk = np.exp(-dtau / FT(q.mu0))
            tsum = ctl.tau_sum[:, iz]
            ier, iet, ier_pm, iet_mm, ieJp, ieJm = elemental_inelastic(model, rrs, m, nd, dtau, fsc[:, iz], tsum, F0,
                                                                      Zpp_r, Zmp_r)
            r, t, r_pm, t_mm, j0p, j0m = orc.elemental(model, m, nd, dtau, p.varpi, p.Zpp, p.Zmp, tsum, F0)
            if nd > 0:
                (r, t, r_pm, t_mm, j0p, j0m, ier, iet, ier_pm, iet_mm, ieJp, ieJm) = doubling_inelastic(
                    model, rrs, nd, expk, r, t, j0p, j0m, ier, iet, ieJp, ieJm)
            if iz == 0:
                C = Composite(r.copy(), r_pm.copy(), t.copy(), t_mm.copy(), j0p.copy(), j0m.copy())
                CI = CompositeIE(ier.copy(), ier_pm.copy(), iet.copy(), iet_mm.copy(), ieJp.copy(), ieJm.copy())
            else:
                C, CI = interaction_inelastic_11(rrs, C, CI, r, r_pm, t, t_mm, j0p, j0m, ier, ier_pm, iet, iet_mm,
                                                 ieJp, ieJm)
        s = orc.brdf_surface(model, m, ctl.tau_sum[:, -1]) if model.surface_rho is not None \
            else orc.lambertian_surface(model, m, ctl.tau_sum[:, -1])
        r_s, rpm_s, t_s, tmm_s, j0p_s, j0m_s = s
        zM = np.zeros((nR, nS, N, N), FT)
        zV = np.zeros((nR, nS, N), FT)
        C, CI = interaction_inelastic_11(rrs, C, CI, np.array(r_s), np.array(rpm_s), np.array(t_s), np.array(tmm_s),
                                         j0p_s, j0m_s, zM, zM, zM, zM, zV, zV)
        for i in range(nV):
            imu = nearest_point(q.qp_mu, FT(cosd(model.vza[i])))
            ist = imu * pol_n
            c = cosd(m * model.vaz[i])
            sn = sind(m * model.vaz[i])
            wv = (weight * np.array([c, c, sn, sn][:pol_n])).astype(FT)
            R_SFI[i] += wv[:, None] * C.J0m[:, ist:ist + pol_n].T
            T_SFI[i] += wv[:, None] * C.J0p[:, ist:ist + pol_n].T
            for dn in range(nR):                                       # postprocessing_vza.jl:143-146
                ieR_SFI[i] += wv[:, None] * CI.ieJ0m[dn, :, ist:ist + pol_n].T
                ieT_SFI[i] += wv[:, None] * CI.ieJ0p[dn, :, ist:ist + pol_n].T
        if return_streams:
            streams.append((C, CI))
    out = dict(R_SFI=R_SFI, T_SFI=T_SFI, ieR_SFI=ieR_SFI, ieT_SFI=ieT_SFI)
    if return_streams:
        out['streams'] = streams
    return out


def rt_run_rrs_points(model: Model, rrs: RRS, idx, F0=None):
    """The RRS result at the band indices ``idx`` only, without running the whole band: a unit (n1, dn) reads nothing
    but the spectral points n1 and n0 = n1 + i_shift[dn], and its contribution is linear in ϖ_λ₁λ₀[dn], so every
    unit is evaluated as the first point of a two-point band [n1, n0] with one Raman line of shift +1 and weight 1
    (the spurious coupling n0 -> next pair only lands on the odd, ignored outputs).  Control flow from the full band.
    Returns dict(R_SFI, T_SFI, ieR_SFI, ieT_SFI) of shape (nVZA, pol_n, len(idx)).  Used by bench.py (parity check
    of the benchmarked band, CPU baseline sample)."""
    import copy
    idx = np.asarray(idx, int)
    nS = model.nSpec
    ctl_full = orc.band_control(model, 0)
    pairs, owner, wts = [], [], []
    for q, n1 in enumerate(idx):
        for dn, sh in enumerate(rrs.i_shift):
            n0 = n1 + int(sh)
            if 0 <= n0 < nS:
                pairs += [n1, n0]
                owner.append(q)
                wts.append(rrs.varpi_raman[dn])
    if not pairs:                                   # no coupling at all: elastic only
        pairs = [v for n1 in idx for v in (n1, n1)]
        owner = list(range(len(idx)))
        wts = [0.0] * len(idx)
    pairs = np.asarray(pairs)
    sub = copy.copy(model)
    sub.tau_rayl = model.tau_rayl[pairs]
    sub.tau_abs = model.tau_abs[pairs]
    Fb = model.F0 if F0 is None else F0
    sub.F0 = None if Fb is None else np.asarray(Fb)[:, pairs]
    out = rt_run_rrs(sub, RRS(np.array([1]), np.array([1.0]), rrs.greek_raman), ctl_override=ctl_full)
    nV, pn = out["R_SFI"].shape[:2]
    res = {k: np.zeros((nV, pn, len(idx)), model.FT) for k in ("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI")}
    seen = set()
    for u, (q, w) in enumerate(zip(owner, wts)):
        res["ieR_SFI"][:, :, q] += w * out["ieR_SFI"][:, :, 2 * u]
        res["ieT_SFI"][:, :, q] += w * out["ieT_SFI"][:, :, 2 * u]
        if q not in seen:
            seen.add(q)
            res["R_SFI"][:, :, q] = out["R_SFI"][:, :, 2 * u]
            res["T_SFI"][:, :, q] = out["T_SFI"][:, :, 2 * u]
    return res


# --------------------------------------------------------------------------
# rt_run_ss(::RRS, model, iBand)   rt_run.jl:593-755 with rt_kernel_ss!(::RRS) CoreKernel/rt_kernel_ss.jl:240-331 (elemental_inelastic!
# and elemental! of the WHOLE layer with ndoubl = 0) and interaction_ss! / get_interaction_ss_RRS! (CoreKernel/interaction_ss.jl:14-19, 70-90)
# --------------------------------------------------------------------------
def rt_run_ss_rrs(model: Model, rrs: RRS):
    """Single-scattering driver with rotational Raman: dict(R_SFI, T_SFI, ieR_SFI, ieT_SFI, hem_R, hem_T)."""
    FT = model.FT
    q = model.quad
    pol_n = model.pol_n
    nS = model.nSpec
    nV = len(model.vza)
    nR = len(rrs.i_shift)
    mu = q.qp_muN.astype(FT)
    el = orc.rt_run_ss(model)
    ieR = np.zeros((nV, pol_n, nS), FT)
    ieT = np.zeros((nV, pol_n, nS), FT)
    F0 = model.F0
    if F0 is None:
        F0 = np.zeros((pol_n, nS), FT)
        F0[0, :] = 1
    fsc = fscatt_rayleigh(model)
    valid = np.zeros((nR, nS), bool)
    for dn in range(nR):
        _, s1 = get_n0_n1(nS, int(rrs.i_shift[dn]))
        valid[dn, s1] = True
    for m in range(model.m_max + 1):
        weight = FT(0.5 / math.pi) if m == 0 else FT(1.0 / math.pi)
        Zpp_r, Zmp_r = orc.compute_Z_moments(pol_n, q.qp_mu, rrs.greek_raman, m, FT)
        ctl = orc.band_control(model, m)
        props = orc.construct_core_optical_properties(model, m)
        Jp = Jm = None
        for iz in range(model.Nz):
            tau = np.broadcast_to(props[iz].tau, (nS,)).astype(FT)
            tsum = ctl.tau_sum[:, iz]
            _, _, _, _, jp, jm = elemental_inelastic(model, rrs, m, 0, tau, fsc[:, iz], tsum, F0, Zpp_r, Zmp_r)
            if iz == 0:
                Jp, Jm = jp.copy(), jm.copy()
            else:                                          # get_interaction_ss_RRS!: attenuation at the OUTPUT wavelength n1
                upd_p = Jp * np.exp(-tau[None, :, None] / mu[None, None, :]) + jp
                upd_m = Jm + jm * np.exp(-tsum[None, :, None] / mu[None, None, :])
                Jp = np.where(valid[:, :, None], upd_p, Jp)
                Jm = np.where(valid[:, :, None], upd_m, Jm)
        for i in range(nV):
            imu = nearest_point(q.qp_mu, FT(cosd(model.vza[i])))
            ist = imu * pol_n
            c = cosd(m * model.vaz[i])
            sn = sind(m * model.vaz[i])
            wv = (weight * np.array([c, c, sn, sn][:pol_n])).astype(FT)
            for dn in range(nR):
                ieR[i] += wv[:, None] * Jm[dn, :, ist:ist + pol_n].T
                ieT[i] += wv[:, None] * Jp[dn, :, ist:ist + pol_n].T
    return dict(R_SFI=el["R_SFI"], T_SFI=el["T_SFI"], ieR_SFI=ieR, ieT_SFI=ieT, hem_R=el["hem_R"], hem_T=el["hem_T"])

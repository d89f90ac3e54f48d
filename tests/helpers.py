"""Shared test helpers: build the same scene for the CUDA path (vsmom_b200.RTModel)
and for the CPU oracle (vsmom_oracle.Model)."""
import json
import math
import os

import numpy as np

import vsmom_oracle as orc
import vsmom_b200 as vb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tables.json")


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def to_oracle(model: vb.RTModel) -> orc.Model:
    q = model.quad_points
    oq = orc.QuadPoints(q.mu0, q.imu0 - 1, q.qp_mu, q.wt_mu, q.qp_muN, q.wt_muN, q.Nquad, q.Nstreams, model.pol_n)

    def g(x):
        return orc.GreekCoefs(np.asarray(x.alpha, float), np.asarray(x.beta, float), np.asarray(x.gamma, float),
                              np.asarray(x.delta, float), np.asarray(x.epsilon, float), np.asarray(x.zeta, float))
    return orc.Model(pol_n=model.pol_n, quad=oq, vza=list(model.vza), vaz=list(model.vaz),
                     tau_rayl=model.tau_rayl, tau_abs=model.tau_abs, tau_aer=model.tau_aer,
                     aerosols=[orc.Aerosol(g(a.greek_coefs), a.omega, a.ft) for a in model.aerosol_optics],
                     greek_rayleigh=g(model.greek_rayleigh), albedo=model.albedo, m_max=model.m_max,
                     varpi_cabannes=model.varpi_cabannes,
                     dtau_max_threshold=model.numerics.dtau_max_threshold,
                     dtau_min_floor=model.numerics.dtau_min_floor, FT=model.float_type,
                     surface_rho=model.surface_rho, ss_correction=model.ss_correction,
                     thermal_B=model.thermal_B, sif0=model.sif0)


def rayleigh_slab_model(pol, nstreams, sza, vza, vaz, tau, albedo, FT=np.float64, nSpec=2, depol=0.0):
    """Single Rayleigh layer with overridden τ, as the reference's Natraj / 6SV1 tests
    build it (test/test_CoreRT.jl:18-26,125-132)."""
    q = vb.rt_set_streams(nstreams, sza, vza, vb.model.POL_N[pol], FT)
    return vb.RTModel(pol, q, sza, list(vza), list(vaz), np.full((nSpec, 1), tau, FT), np.zeros((nSpec, 1), FT),
                      np.zeros((0, 1)), [], vb.get_greek_rayleigh(depol), albedo, 2, 1.0,
                      vb.RTNumericalParameters(), FT)


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a, float) - np.asarray(b, float))) / max(np.max(np.abs(np.asarray(b, float))), 1e-300))


def cds_golden():
    """Subset of the Natraj et al. (2009) CDS tables (tests/golden/natraj_cds_subset.json)."""
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "natraj_cds_subset.json")))


def cds_model(d, albedo, mu0):
    """The model of test/benchmarks/benchmark_natraj_cds.jl:242-256: one Rayleigh layer, Stokes_IQUV, 11 Gauss
    streams, all 16 x 7 (mu, phi) of the tables as viewing geometries, the optical depths as spectral axis."""
    import math
    vza = [math.degrees(math.acos(m)) for _ in d["phi"] for m in d["mu"]]
    vaz = [p for p in d["phi"] for _ in d["mu"]]
    sza = math.degrees(math.acos(mu0))
    q = vb.rt_set_streams(d["nstreams"], sza, vza, 4, np.float64)
    nS = len(d["tau"])
    return vb.RTModel("Stokes_IQUV", q, sza, vza, vaz, np.array(d["tau"], float)[:, None], np.zeros((nS, 1)),
                      np.zeros((0, 1)), [], vb.get_greek_rayleigh(d["depol"]), albedo, 2, 1.0,
                      vb.RTNumericalParameters(), np.float64)


def cds_errors(d, out, ia, i0):
    """Worst absolute / relative (|truth| >= 0.01) deviation of pi*R_SFI, pi*T_SFI from the tables per Stokes
    component and direction; downwelling I without the direct beam at the solar ordinate
    (benchmark_natraj_cds.jl:262-272)."""
    import math
    mus, phis, mu0 = d["mu"], d["phi"], d["mu0"][i0]
    worst = {}
    for di, arr in (("UP", math.pi * out["R_SFI"]), ("DN", math.pi * out["T_SFI"])):
        for si, st in enumerate("IQU"):
            for it, tau in enumerate(d["tau"]):
                tab = np.array(d["tables"][f"{st}_{di}"][it][ia][i0])          # [mu][phi]
                mod = arr[:, si, it].reshape(len(phis), len(mus)).T.copy()     # observation index = iphi*16 + imu
                if di == "DN" and st == "I":
                    for imu, mu in enumerate(mus):
                        if abs(mu - mu0) < 1e-12:
                            mod[imu, :] -= 0.5 * math.exp(-tau / mu0)
                ae = np.abs(mod - tab)
                big = np.abs(tab) >= 0.01
                rel = float((ae[big] / np.abs(tab[big])).max()) if big.any() else 0.0
                w = worst.get(st, (0.0, 0.0))
                worst[st] = (max(w[0], float(ae.max())), max(w[1], rel))
    return worst


# measured over the whole subset: abs <= 9.2e-5; rel I 3.7e-4, Q 3.3e-3, U 1.8e-4 (the reference's own Natraj
# limits at tau = 0.5, mu0 = 0.2 are 5e-4 / 2.5e-3 / 5e-4, test/test_CoreRT.jl:153-155)
CDS_TOL = {"abs": 1.5e-4, "I": 6e-4, "Q": 5e-3, "U": 6e-4}


def solar_tester_model(st, sza, raz):
    """VLIDORT 2.8.3 solar_tester scene as test/vlidort_baseline/cases/case_B_solar_tester.jl builds it: 23 layers
    (tau_rayl = molomg*molext, tau_abs = (1-molomg)*molext), HG aerosol g = 0.8, omega = 0.95, total tau 0.5 spread
    over the bottom 6 layers by geometric thickness, 16 phase moments, Stokes_I, 8 Gauss streams, Lambertian 0.05,
    two identical spectral points."""
    a = st["aerosol"]
    L = a["nmoments"]
    beta = np.array([(2 * l + 1) * a["g"] ** l for l in range(L + 1)])
    z = np.zeros(L + 1)
    aer = vb.AerosolOptics(vb.GreekCoefs(z.copy(), beta, z.copy(), z.copy(), z.copy(), z.copy()), a["omega"], 0.0)
    h = np.array([60.0] + list(st["height_km"]))
    n6 = 23 - a["layers"]
    parcel = a["tau_total"] / (h[n6] - h[-1])
    aerext = np.zeros(23)
    aerext[n6:] = parcel * (h[n6:-1] - h[n6 + 1:])
    ext, ssa = np.array(st["molext"]), np.array(st["molomg"])
    q = vb.rt_set_streams(st["nstreams"], sza, st["vza"], 1, np.float64)
    return vb.RTModel("Stokes_I", q, sza, list(st["vza"]), [raz] * len(st["vza"]), np.tile(ssa * ext, (2, 1)),
                      np.tile((1 - ssa) * ext, (2, 1)), aerext[None, :], [aer], vb.get_greek_rayleigh(st["depol"]),
                      st["albedo"], vb.derive_m_max([aer], st["nstreams"]), 1.0, vb.RTNumericalParameters(), np.float64)


def solar_tester_errors(st, out, isza, iraz):
    """Worst relative deviation of R_SFI (TOA up) and T_SFI (BOA down) from VLIDORT task 1 over the VZAs."""
    up = dn = 0.0
    for ivza in range(len(st["vza"])):
        g = (isza * 3 + ivza) * 3 + iraz
        up = max(up, abs(out["R_SFI"][ivza, 0, 0] - st["toa_up"][g]) / st["toa_up"][g])
        dn = max(dn, abs(out["T_SFI"][ivza, 0, 0] - st["boa_dn"][g]) / st["boa_dn"][g])
    return up, dn


def solar_tester_vector_model(st, sv, sza, raz):
    """The NSTOKES = 3 variant (test/vlidort_baseline/cases/case_C_solar_tester_vector.jl): same 23-layer atmosphere,
    Problem III Mie aerosol (first 16 F-matrix moments: a1->beta, b1->gamma, a2->alpha, a3->zeta, b2->-epsilon,
    a4->delta), omega 0.99999, Stokes_IQU."""
    mo = sv["moms"]
    g = vb.GreekCoefs(np.array(mo["a2"]), np.array(mo["a1"]), np.array(mo["b1"]), np.array(mo["a4"]),
                      -np.array(mo["b2"]), np.array(mo["a3"]))
    aer = vb.AerosolOptics(g, sv["omega"], 0.0)
    h = np.array([60.0] + list(st["height_km"]))
    n6 = 23 - sv["layers"]
    parcel = sv["tau_total"] / (h[n6] - h[-1])
    aerext = np.zeros(23)
    aerext[n6:] = parcel * (h[n6:-1] - h[n6 + 1:])
    ext, ssa = np.array(st["molext"]), np.array(st["molomg"])
    q = vb.rt_set_streams(sv["nstreams"], sza, sv["vza"], 3, np.float64)
    return vb.RTModel("Stokes_IQU", q, sza, list(sv["vza"]), [raz] * len(sv["vza"]), np.tile(ssa * ext, (2, 1)),
                      np.tile((1 - ssa) * ext, (2, 1)), aerext[None, :], [aer], vb.get_greek_rayleigh(st["depol"]),
                      sv["albedo"], vb.derive_m_max([aer], sv["nstreams"]), 1.0, vb.RTNumericalParameters(), np.float64)


def solar_tester_vector_errors(sv, out, isza, iraz):
    """Worst relative deviation per direction over I/Q/U and the VZAs; VLIDORT's Q and U carry the opposite sign
    (case_C_solar_tester_vector.jl:163-166)."""
    up = dn = 0.0
    for si, s in enumerate("IQU"):
        sgn = 1.0 if s == "I" else -1.0
        for ivza in range(len(sv["vza"])):
            g = (isza * 3 + ivza) * 3 + iraz
            tu, td = sgn * sv[s + "_toa_up"][g], sgn * sv[s + "_boa_dn"][g]
            up = max(up, abs(out["R_SFI"][ivza, si, 0] - tu) / abs(tu))
            dn = max(dn, abs(out["T_SFI"][ivza, si, 0] - td) / abs(td))
    return up, dn

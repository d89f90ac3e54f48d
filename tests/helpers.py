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
                     surface_rho=model.surface_rho)


def rayleigh_slab_model(pol, nstreams, sza, vza, vaz, tau, albedo, FT=np.float64, nSpec=2, depol=0.0):
    """Single Rayleigh layer with overridden τ, as the reference's Natraj / 6SV1 tests
    build it (test/test_CoreRT.jl:18-26,125-132)."""
    q = vb.rt_set_streams(nstreams, sza, vza, vb.model.POL_N[pol], FT)
    return vb.RTModel(pol, q, sza, list(vza), list(vaz), np.full((nSpec, 1), tau, FT), np.zeros((nSpec, 1), FT),
                      np.zeros((0, 1)), [], vb.get_greek_rayleigh(depol), albedo, 2, 1.0,
                      vb.RTNumericalParameters(), FT)


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a, float) - np.asarray(b, float))) / max(np.max(np.abs(np.asarray(b, float))), 1e-300))

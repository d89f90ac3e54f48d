"""LinMode at N = 45 (config-2 shape, 34 layers, P = 4): zero-padded on the N = 60 streamed kernels vs the run-time-N SIMT kernels."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("vsmartmom.jl_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr
from test_gpu_lin import lin_params
m = vb.synthetic_o2a_model(nSpec=148, nstreams=10, polarization_type="Stokes_IQU", aerosol=True)
for a in m.aerosol_optics:
    g = a.greek_coefs
    for f in ("alpha", "beta", "gamma", "delta", "epsilon", "zeta"):
        setattr(g, f, getattr(g, f)[:6])
m.m_max = 5
d = lin_params(m)
out = {}
for mode in ("embedded", "generic"):
    if mode == "generic": os.environ["VSM_DISABLE_EMBED"] = "1"
    else: os.environ.pop("VSM_DISABLE_EMBED", None)
    ctx = _lib.Context(0)
    for _ in range(2):
        r = rr.rt_run_lin(m, *d, ctx=ctx)
    out[mode] = dict(ms=ctx.last_column_ms()[0], dR=r["dR_SFI"])
    ctx.close()
e = float(np.max(np.abs(out["embedded"]["dR"] - out["generic"]["dR"])) / np.max(np.abs(out["generic"]["dR"])))
print(json.dumps(dict(N=45, nSpec=148, P=4, n_moments=6, embedded_ms=out["embedded"]["ms"], generic_ms=out["generic"]["ms"],
                      speedup=out["generic"]["ms"] / out["embedded"]["ms"], max_rel_dR=e)))

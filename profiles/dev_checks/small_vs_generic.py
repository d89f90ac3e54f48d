"""Device time of the warp-per-point small-matrix kernel (vsm_small.cuh, N <= 16) against the CTA-per-item generic
kernels on the same scenes; also checks that both paths agree.  Run on a B200:  python profiles/dev_checks/small_vs_generic.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vsmartmom.jl_b200"))
import numpy as np
import torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr

scenes = {
    "quickstart-like: Stokes_I, 3 streams, 2 layers, Rayleigh, 3 moments": dict(nSpec=20000, nstreams=3, polarization_type="Stokes_I", aerosol=False, vza=[60.0], vaz=[0.0], p_half=np.array([0.1, 500.0, 1013.0])),
    "file default: Stokes_I, 8 streams, 34 layers, Rayleigh + aerosol, 16 moments": dict(nSpec=10000, nstreams=8, polarization_type="Stokes_I", aerosol=True),
    "Stokes_IQU, 3 streams, 34 layers, Rayleigh": dict(nSpec=10000, nstreams=3, polarization_type="Stokes_IQU", aerosol=False, vza=[0.0], vaz=[0.0]),
}
out = {}
for name, kw in scenes.items():
    m = vb.synthetic_o2a_model(**kw)
    res = {}
    for mode in ("small", "generic"):
        if mode == "generic":
            os.environ["VSM_DISABLE_SMALL"] = "1"
        else:
            os.environ.pop("VSM_DISABLE_SMALL", None)
        ctx = _lib.Context(0)
        prep = rr.prepare(m, "cuda:0")
        tr = torch.from_numpy(np.ascontiguousarray(m.tau_rayl.T)).cuda()
        ta = torch.from_numpy(np.ascontiguousarray(m.tau_abs.T)).cuda()
        f0 = np.zeros((m.nSpec, m.pol_n)); f0[:, 0] = 1
        f0 = torch.from_numpy(f0).cuda()
        ms = []
        for _ in range(4):
            R, T, H, uw, dw = rr.rt_column_device(ctx, prep, tr, ta, f0)
            ctx.synchronize()
            ms.append(ctx.last_column_ms()[0])
        res[mode] = dict(ms=min(ms[1:]), R=R.cpu().numpy(), T=T.cpu().numpy(), H=H.cpu().numpy(), uw=uw.cpu().numpy())
        ctx.close()
    err = max(float(np.max(np.abs(res["small"][k] - res["generic"][k])) / np.max(np.abs(res["generic"][k]))) for k in ("R", "T", "H", "uw"))
    out[name] = dict(N=int(m.quad_points.qp_muN.size), nSpec=m.nSpec, small_ms=res["small"]["ms"], generic_ms=res["generic"]["ms"],
                     speedup=res["generic"]["ms"] / res["small"]["ms"], points_per_s_small=m.nSpec / res["small"]["ms"] * 1e3,
                     max_rel_small_vs_generic=err)
print(json.dumps(out, indent=1))

"""Device time of a column whose N has no specialised kernel, run (a) zero-padded as the next specialised size and
(b) on the run-time-N generic kernels (VSM_DISABLE_EMBED=1).  Run on a B200."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vsmartmom.jl_b200"))
import numpy as np
import torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr

out = {}
for name, kw, nS in (("file default with Stokes_IQU: 8 streams -> N = 39 (runs as 45)", dict(nstreams=8, polarization_type="Stokes_IQU"), 2000),
                     ("Stokes_IQU, 4 streams -> N = 27 (runs as 45)", dict(nstreams=4, polarization_type="Stokes_IQU"), 2000),
                     ("Stokes_IQUV, 12 streams -> N = 68 (runs as 75)", dict(nstreams=12, polarization_type="Stokes_IQUV"), 296)):
    m = vb.synthetic_o2a_model(nSpec=nS, aerosol=True, **kw)
    res = {}
    for mode in ("embedded", "generic"):
        if mode == "generic":
            os.environ["VSM_DISABLE_EMBED"] = "1"
        else:
            os.environ.pop("VSM_DISABLE_EMBED", None)
        ctx = _lib.Context(0)
        prep = rr.prepare(m, "cuda:0")
        tr = torch.from_numpy(np.ascontiguousarray(m.tau_rayl.T)).cuda(); ta = torch.from_numpy(np.ascontiguousarray(m.tau_abs.T)).cuda()
        f0 = np.zeros((m.nSpec, m.pol_n)); f0[:, 0] = 1
        f0 = torch.from_numpy(f0).cuda()
        ms = []
        for _ in range(3):
            R, T, H, uw, dw = rr.rt_column_device(ctx, prep, tr, ta, f0)
            ctx.synchronize()
            ms.append(ctx.last_column_ms()[0])
        res[mode] = dict(ms=min(ms[1:]), R=R.cpu().numpy())
        ctx.close()
    out[name] = dict(N=int(m.quad_points.qp_muN.size), nSpec=nS, n_moments=m.m_max + 1, embedded_ms=res["embedded"]["ms"], generic_ms=res["generic"]["ms"],
                     speedup=res["generic"]["ms"] / res["embedded"]["ms"], points_per_s=nS / res["embedded"]["ms"] * 1e3,
                     max_rel=float(np.max(np.abs(res["embedded"]["R"] - res["generic"]["R"])) / np.max(np.abs(res["generic"]["R"]))))
print(json.dumps(out, indent=1))

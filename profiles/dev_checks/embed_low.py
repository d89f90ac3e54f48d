import json, os, sys
sys.path.insert(0, "vsmartmom.jl_b200")
import numpy as np, torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr
for ns in (12, 14, 16):
    m = vb.synthetic_o2a_model(nSpec=2000, nstreams=ns, polarization_type="Stokes_I", aerosol=True)
    res = {}
    for mode, env in (("embedded", {"VSM_EMBED_MIN": "17"}), ("generic", {"VSM_DISABLE_EMBED": "1"})):
        for k in ("VSM_EMBED_MIN", "VSM_DISABLE_EMBED"): os.environ.pop(k, None)
        os.environ.update(env)
        ctx = _lib.Context(0); prep = rr.prepare(m, "cuda:0")
        tr = torch.from_numpy(np.ascontiguousarray(m.tau_rayl.T)).cuda(); ta = torch.from_numpy(np.ascontiguousarray(m.tau_abs.T)).cuda()
        f0 = np.zeros((m.nSpec, m.pol_n)); f0[:, 0] = 1; f0 = torch.from_numpy(f0).cuda()
        ms = []
        for _ in range(3):
            R, *_ = rr.rt_column_device(ctx, prep, tr, ta, f0); ctx.synchronize(); ms.append(ctx.last_column_ms()[0])
        res[mode] = (min(ms[1:]), R.cpu().numpy()); ctx.close()
    print(m.quad_points.qp_muN.size, res["embedded"][0], res["generic"][0], res["generic"][0] / res["embedded"][0],
          float(np.max(np.abs(res["embedded"][1] - res["generic"][1])) / np.max(np.abs(res["generic"][1]))))

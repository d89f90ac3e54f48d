"""Device time of one Raman owner shard (1/8 of the band, balanced by units) on ONE GPU: total / inelastic / elastic ms."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vsmartmom.jl_b200"))
import numpy as np, torch
from vsmom_b200 import _lib, raman, rt_run as rr
m = raman.synthetic_rrs_model(); rs = raman.synthetic_rrs()
ctx = _lib.Context(0)
prep = rr.prepare(m, "cuda:0")
tr = torch.from_numpy(np.ascontiguousarray(m.tau_rayl.T)).cuda(); ta = torch.from_numpy(np.ascontiguousarray(m.tau_abs.T)).cuda()
f0 = np.zeros((m.nSpec, m.pol_n)); f0[:, 0] = 1; f0 = torch.from_numpy(f0).cuda()
zp, zm = raman.raman_z_tables(m, rs); zp = torch.from_numpy(zp).cuda(); zm = torch.from_numpy(zm).cuda()
out = {}
for world in (1, 8):
    b = raman.rrs_shard_bounds(m.nSpec, rs, world)
    for r in sorted({0, world // 2}):
        lo, hi = b[r]
        for _ in range(3):
            raman.rrs_column_device(ctx, prep, rs, tr, ta, f0, zp, zm, lo, hi - lo); ctx.synchronize()
        out[f"world{world}_rank{r}"] = dict(owner=(lo, hi), ms=ctx.last_column_ms())
for order in ("0",):
    os.environ["VSM_RRS_ORDER"] = order
    c2 = _lib.Context(0)
    lo, hi = raman.rrs_shard_bounds(m.nSpec, rs, 8)[4]
    for _ in range(3):
        raman.rrs_column_device(c2, prep, rs, tr, ta, f0, zp, zm, lo, hi - lo); c2.synchronize()
    out["world8_rank4_order0"] = dict(owner=(lo, hi), ms=c2.last_column_ms())
print(json.dumps(out, indent=1))

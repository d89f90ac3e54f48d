"""One column call of the file-default-shaped scene (Stokes_I, 8 streams -> N = 13, 34 layers, 16 moments) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vsmartmom.jl_b200"))
import numpy as np, torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr
m = vb.synthetic_o2a_model(nSpec=int(sys.argv[1]) if len(sys.argv) > 1 else 4000, nstreams=8, polarization_type="Stokes_I", aerosol=True)
ctx = _lib.Context(0); prep = rr.prepare(m, "cuda:0")
tr = torch.from_numpy(np.ascontiguousarray(m.tau_rayl.T)).cuda(); ta = torch.from_numpy(np.ascontiguousarray(m.tau_abs.T)).cuda()
f0 = np.zeros((m.nSpec, m.pol_n)); f0[:, 0] = 1; f0 = torch.from_numpy(f0).cuda()
for _ in range(2):
    rr.rt_column_device(ctx, prep, tr, ta, f0); ctx.synchronize()
print(ctx.last_column_ms())

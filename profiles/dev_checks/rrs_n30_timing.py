import os, sys, json, time
sys.path.insert(0, "vsmartmom.jl_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import numpy as np
from vsmom_b200 import _lib, raman
import test_gpu_rrs as t
m, rs = t._scene(nstreams=8, nS=400, Nz=3, aerosol=True)
ctx = _lib.Context(0)
for _ in range(2):
    raman.rt_run_rrs(m, rs, ctx=ctx)
print(json.dumps(dict(N=int(m.quad_points.qp_muN.size), nSpec=400, n_raman=8, ms=ctx.last_column_ms())))

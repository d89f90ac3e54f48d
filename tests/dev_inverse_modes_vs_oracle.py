"""Developer check: errors of the two inverse modes against the oracle (run on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vsmartmom.jl_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import vsmom_b200 as vb, vsmom_oracle as orc
from vsmom_b200 import _lib, rt_run as rr
from helpers import to_oracle, relerr
ctx = _lib.Context(0)
for scale in (1.0, 8.0, 40.0):
    m = vb.synthetic_o2a_model(nSpec=12, aerosol=True, albedo=0.8)
    m.tau_rayl = m.tau_rayl * scale; m.tau_aer = m.tau_aer * scale; m.tau_abs = m.tau_abs * np.logspace(-5.0, -0.5, 12)[:, None]
    ref = orc.rt_run(to_oracle(m))
    ctx.set_inverse_mode(0); gj = rr.rt_run(m, ctx)
    ctx.set_inverse_mode(1); ad = rr.rt_run(m, ctx)
    for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw"):
        print(f"scale {scale:5.1f} {k:7s} gj-ref {relerr(gj[k], ref[k]):.2e}  ad-ref {relerr(ad[k], ref[k]):.2e}  ad-gj {relerr(ad[k], gj[k]):.2e}  max|ref| {np.abs(ref[k]).max():.2e}")

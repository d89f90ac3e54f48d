"""Developer script: quick GPU-vs-oracle parity + timing on a few configurations."""
import sys, os, time, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vsmartmom.jl_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr
import vsmom_oracle as orc
from helpers import to_oracle, rayleigh_slab_model, relerr, golden

ctx = _lib.Context(0)
def check(name, model, nref=None):
    t0 = time.time(); out = rr.rt_run(model, ctx); t1 = time.time()
    sel = slice(None) if nref is None else slice(0, nref)
    ref = orc.rt_run(to_oracle(model), sel); t2 = time.time()
    errs = {k: relerr(out[k][..., sel] if out[k].ndim == 3 else out[k][sel], ref[k]) for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw")}
    print(f"{name}: N={model.quad_points.qp_muN.size} nSpec={model.nSpec} Nz={model.Nz} m_max={model.m_max} FT={model.float_type.__name__} gpu={t1-t0:.3f}s cpu={t2-t1:.3f}s ms={ctx.last_column_ms()} errs={errs}", flush=True)

which = sys.argv[1:] or ["tiny", "o2a", "o2a32", "natraj"]
if "tiny" in which:
    m = vb.synthetic_o2a_model(nSpec=16, nstreams=3, polarization_type="Stokes_I", aerosol=False, vza=[60.0], vaz=[180.0], p_half=np.array([100.0, 500.0, 1000.0]))
    check("quickstart-like", m)
if "o2a" in which:
    check("o2a-rayleigh", vb.synthetic_o2a_model(nSpec=64, aerosol=False))
    check("o2a-aerosol", vb.synthetic_o2a_model(nSpec=64, aerosol=True), nref=16)
if "o2a32" in which:
    check("o2a-aerosol-f32", vb.synthetic_o2a_model(nSpec=64, aerosol=True, float_type=np.float32), nref=16)
if "natraj" in which:
    d = golden()["natraj"]
    vza = [math.degrees(math.acos(x)) for x in d["mu"]]
    check("natraj", rayleigh_slab_model("Stokes_IQUV", 11, math.degrees(math.acos(0.2)), vza, [30.0] * 16, 0.5, 0.0))
if "big" in which:
    m = vb.synthetic_o2a_model(nSpec=2000, aerosol=True)
    prep = rr.prepare(m)
    for _ in range(2):
        t0 = time.time(); out = rr.rt_run(m, ctx, prep); t1 = time.time()
        print("big", t1 - t0, ctx.last_column_ms(), flush=True)

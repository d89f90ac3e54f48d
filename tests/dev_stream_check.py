"""Developer check (not collected by pytest): streamed kernels (N = 64/75/100) vs the oracle and vs the generic path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "vsmartmom.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import vsmom_b200 as vb
import vsmom_oracle as orc
from vsmom_b200 import _lib, rt_run as rr
from helpers import to_oracle, relerr

def run(pol, nstreams, nspec, mmax=None, noscat=False):
    m = vb.synthetic_ocean_model(nSpec=nspec, nstreams=nstreams, polarization_type=pol)
    if mmax is not None:
        m.m_max = mmax; m.surface_rho = m.surface_rho[:mmax + 1]
        for a in m.aerosol_optics:
            g = a.greek_coefs
            for f in ("alpha", "beta", "gamma", "delta", "epsilon", "zeta"):
                setattr(g, f, getattr(g, f)[:mmax + 1])
    if noscat:
        m.tau_rayl[:, 0] = 0; m.tau_rayl[:, 2] = 0
    N = m.quad_points.qp_muN.size
    ctx = _lib.Context(0)
    t = time.time(); out = rr.rt_run(m, ctx); t1 = time.time() - t
    t = time.time(); out = rr.rt_run(m, ctx); t2 = time.time() - t
    ms = ctx.last_column_ms()
    ctx.close()
    line = f"{pol} nstreams={nstreams} N={N} nSpec={nspec} moments={m.m_max+1}: stream ms={ms} "
    if nspec <= 4 or os.environ.get("DEV_GENERIC"):
        os.environ["VSM_DISABLE_STREAM"] = "1"
        ctx = _lib.Context(0)
        gen = rr.rt_run(m, ctx); msg = ctx.last_column_ms(); ctx.close()
        del os.environ["VSM_DISABLE_STREAM"]
        line += f"generic ms={msg} "
        for k in ("R_SFI", "T_SFI", "hdr"):
            line += f"| {k} vs generic {relerr(out[k], gen[k]):.2e} "
    print(line, flush=True)
    if nspec <= 4:
        t = time.time(); ref = orc.rt_run(to_oracle(m)); t3 = time.time() - t
        print("   vs oracle:", {k: f"{relerr(out[k], ref[k]):.2e}" for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw")}, f"(oracle {t3:.0f}s)", flush=True)

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    if what == "small":
        run("Stokes_IQU", 20, 2, mmax=3)
        run("Stokes_IQUV", 20, 2, mmax=3)
        run("Stokes_IQUV", 11, 2, mmax=3)
        run("Stokes_IQU", 20, 2, mmax=2, noscat=True)
    elif what == "prof":
        only = os.environ.get("DEV_ONLY", "")
        if only in ("", "100"):
            run("Stokes_IQUV", 20, 148, mmax=0)
        if only in ("", "75"):
            run("Stokes_IQU", 20, 296, mmax=0)
    else:
        mm = int(os.environ.get("DEV_MMAX", "9"))
        run("Stokes_IQU", 20, 296, mmax=mm)
        run("Stokes_IQUV", 20, 148, mmax=mm)
        run("Stokes_IQUV", 11, 296, mmax=mm)

"""Developer tool: per-phase cycle split of the fused layer kernel (debug build with
-DVSM_PHASE_TIMING; see csrc/vsm_column.cuh).  Run on the GPU box:
    VSMOM_B200_LIB=vsmartmom.jl_b200/lib/libvsmom_b200_dbg.so python profiles/phase_timing.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vsmartmom.jl_b200"))
import numpy as np, torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr
m = vb.synthetic_o2a_model(nSpec=444, aerosol=True)
ctx = _lib.Context(0)
prep = rr.prepare(m)
rr.rt_run(m, ctx, prep)
buf = (C.c_ulonglong * 56)()
_lib.lib.vsm_debug_phase_cycles(buf, 1)
rr.rt_run(m, ctx, prep)
_lib.lib.vsm_debug_phase_cycles(buf, 0)
names = ["elemental", "(unused)", "(I - r r)^-1", "mm t*G", "mv2 x2", "mm r*t", "mm r+=tt*Y", "mm tt*t", "doubling loop total", "store"]
v = np.array(list(buf)[:10], dtype=np.float64)
nd = prep.ctl.ndoubl.sum() * 20 * 444
items = 34 * 20 * 444
print("doublings", nd, "items", items, "ms", ctx.last_column_ms())
for k, nme in enumerate(names):
    per = v[k] / (nd if 1 <= k <= 7 else items)
    print(f"{nme:22s} total {v[k]:.3e} cyc   per {'doubling' if 1 <= k <= 7 else 'item'}: {per:9.0f} cyc")

ev = np.array(list(buf)[10:13], dtype=np.float64) / items
print("elemental split (thread 0 cycles per item): per-stream exp/expm1 + barrier %.0f | element loop %.0f | source vectors %.0f" % tuple(ev))
inames = ["load tiles + sync", "a: stage + pivot block", "a: barrier wait", "b: panels", "b: barrier wait", "c: trailing update"]
for lbl, off, calls in (("k_layer", 16, nd), ("k_adding2 (per group)", 40, 2 * 33 * 20 * 444)):
    iv = np.array(list(buf)[off:off + 6], dtype=np.float64) / max(calls, 1)
    print(f"blocked inverse in {lbl}: thread-0 cycles per call, total {iv.sum():.0f}")
    for k, nme in enumerate(inames):
        print(f"  {nme:26s} {iv[k]:9.0f}")

av = np.array(list(buf)[24:40], dtype=np.float64)
steps11 = 33 * 20 * 444
an = ["(I - A)^-1 (+ vector)", "T01 / T21 product", "2nd product(s) + sync", "last product(s)", "wait for the other group",
      "issue prefetch of next layer", "whole step (call)", "cp.async wait + barrier"]
print("k_adding2 per layer step (cycles; thread 0 of group 0 | group 1):")
for k, nme in enumerate(an):
    print(f"  {nme:30s} {av[k] / steps11:9.0f} {av[8 + k] / steps11:9.0f}")

print("inverse calls", buf[16 + 6] + buf[40 + 6], "pivoted fallbacks", buf[16 + 7] + buf[40 + 7])
print("(I - A)^-1 routes  [K=2, K=3, K=4, K=8 | Gauss-Jordan]:  k_layer", list(buf[48:52]), buf[16 + 6], " k_adding2", list(buf[52:56]), buf[40 + 6])

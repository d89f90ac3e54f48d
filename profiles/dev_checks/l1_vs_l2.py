"""L1 (vsm_rt_layer_*: one layer, one moment per call, composite through HBM) against L2 (vsm_rt_column_*) at the
config-2 shape (N = 45, 34 layers): device time per layer call, and agreement of the end-of-column composite-derived
outputs.  Run on a B200."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("vsmartmom.jl_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, layers, rt_run as rr
import vsmom_oracle as orc
from helpers import to_oracle

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
m = vb.synthetic_o2a_model(nSpec=nS, nstreams=10, polarization_type="Stokes_IQU", aerosol=True)
om = to_oracle(m)
q = m.quad_points
N = q.qp_muN.size
ctx = _lib.Context(0)
dev = "cuda"
mu = torch.from_numpy(q.qp_muN).to(dev); wt = torch.from_numpy(q.wt_muN).to(dev)
F0 = torch.zeros((nS, m.pol_n), dtype=torch.float64, device=dev); F0[:, 0] = 1
mom = 1
ctl = orc.band_control(om, mom)
props = orc.construct_core_optical_properties(om, mom)
ins = []
for iz in range(m.Nz):
    p = props[iz]
    Zpp = np.broadcast_to(p.Zpp, (nS, N, N)) if p.Zpp.ndim == 2 else p.Zpp
    Zmp = np.broadcast_to(p.Zmp, (nS, N, N)) if p.Zmp.ndim == 2 else p.Zmp
    tz = lambda Z: torch.from_numpy(np.ascontiguousarray(Z.transpose(0, 2, 1))).to(dev)
    ins.append((torch.from_numpy(np.ascontiguousarray(p.tau)).to(dev),
                torch.from_numpy(np.ascontiguousarray(np.broadcast_to(p.varpi, p.tau.shape))).to(dev), tz(Zpp), tz(Zmp),
                torch.from_numpy(np.ascontiguousarray(ctl.tau_sum[:, iz])).to(dev)))
res = {}
for rep in range(3):
    comp = layers.make_composite_layer(N, nS, torch.float64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    t0 = time.perf_counter()
    for iz in range(m.Nz):
        tau, varpi, Zpp, Zmp, ts = ins[iz]
        layers.rt_kernel_(ctx, m.pol_n, None, comp, tau, varpi, Zpp, Zmp, ctl.interface[iz], ts, mom, q, F0, iz + 1,
                          ctl.ndoubl[iz], ctl.scatter[iz], mu, wt)
    ctx.synchronize()
    wall = time.perf_counter() - t0
res["L1_ms_per_layer_call"] = 1e3 * wall / m.Nz
res["L1_ms_column_one_moment"] = 1e3 * wall
# L2: all moments
prep = rr.prepare(m, dev)
tr = torch.from_numpy(np.ascontiguousarray(m.tau_rayl.T)).to(dev); ta = torch.from_numpy(np.ascontiguousarray(m.tau_abs.T)).to(dev)
for _ in range(2):
    rr.rt_column_device(ctx, prep, tr, ta, F0, want_hdr=False); ctx.synchronize()
tot = ctx.last_column_ms()[0]
nM = m.m_max + 1
res["L2_ms_column_one_moment"] = tot / nM
res["L2_ms_per_layer"] = tot / nM / (m.Nz + 1)
res["L1_over_L2"] = res["L1_ms_column_one_moment"] / res["L2_ms_column_one_moment"]
# parity of the L1 composite vs the oracle on a few points
idx = np.array([0, nS // 2, nS - 1])
import copy
om2 = copy.copy(om); om2.tau_rayl = om.tau_rayl[idx]; om2.tau_abs = om.tau_abs[idx]
props2 = orc.construct_core_optical_properties(om2, mom)
ctl2 = orc.BandControl(ctl.scatter, ctl.ndoubl, ctl.interface, ctl.tau_sum[idx])
C = None
F0o = np.zeros((m.pol_n, 3)); F0o[0] = 1
for iz in range(m.Nz):
    C = orc.rt_kernel_layer(om2, mom, iz, ctl2, props2[iz], C, F0o, ctl2.tau_sum[:, iz])
got = comp.Rmp.cpu().numpy().transpose(0, 2, 1)[idx]
res["L1_Rmp_max_rel_vs_oracle"] = float(np.max(np.abs(got - C.Rmp)) / np.max(np.abs(C.Rmp)))
res["L1_J0m_max_rel_vs_oracle"] = float(np.max(np.abs(comp.J0m.cpu().numpy()[idx] - C.J0m)) / np.max(np.abs(C.J0m)))
res["N"], res["nSpec"], res["Nz"] = int(N), nS, m.Nz
print(json.dumps(res, indent=1))

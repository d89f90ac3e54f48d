"""Developer check: where does the end-to-end step spend time beyond the device-timed step?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vsmartmom.jl_b200"))
import numpy as np, torch
import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr
nS = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dev = torch.device("cuda", 0)
full = vb.synthetic_o2a_model(nSpec=nS, aerosol=True)
prep = rr.prepare(full, dev)
tr_h = torch.from_numpy(np.ascontiguousarray(full.tau_rayl.T)).pin_memory()
ta_h = torch.from_numpy(np.ascontiguousarray(full.tau_abs.T)).pin_memory()
f0 = np.zeros((nS, 3)); f0[:, 0] = 1
f0_h = torch.from_numpy(f0).pin_memory()
ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
tr_d, ta_d, f0_d = tr_h.to(dev), ta_h.to(dev), f0_h.to(dev)
R_h = torch.empty((nS, 3, len(full.vza)), dtype=torch.float64).pin_memory()
def sync(): torch.cuda.synchronize()
for _ in range(2):
    rr.rt_column_device(ctx, prep, tr_d, ta_d, f0_d, want_hdr=False)
sync()
import gc
print('cpus', len(os.sched_getaffinity(0)), 'loadavg', os.getloadavg())
for rep in range(8):
    if rep == 4: gc.disable(); print('gc disabled')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    R, T, *_ = rr.rt_column_device(ctx, prep, tr_d, ta_d, f0_d, want_hdr=False)
    t1 = time.perf_counter(); e1.record(); sync(); t2 = time.perf_counter()
    print(f"resident: call returned after {1e3*(t1-t0):8.1f} ms, done after {1e3*(t2-t0):8.1f} ms, events {e0.elapsed_time(e1):8.1f} ms, lib {ctx.last_column_ms()}")
for rep in range(6):
    t0 = time.perf_counter()
    a = tr_h.to(dev, non_blocking=True); b = ta_h.to(dev, non_blocking=True); c = f0_h.to(dev, non_blocking=True)
    t1 = time.perf_counter()
    R, T, *_ = rr.rt_column_device(ctx, prep, a, b, c, want_hdr=False)
    t2 = time.perf_counter()
    R_h.copy_(R, non_blocking=True); sync(); t3 = time.perf_counter()
    print(f"e2e: h2d issue {1e3*(t1-t0):6.1f} ms, call {1e3*(t2-t1):8.1f} ms, total {1e3*(t3-t0):8.1f} ms, lib {ctx.last_column_ms()}")

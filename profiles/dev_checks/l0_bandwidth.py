"""Developer check: achieved HBM bandwidth of the L0 batched algebra (vsm_batched_mul / vsm_batch_inv)
at the hot-path shape (N=45, one slice per spectral point).  Algorithmic bytes: mul 3 N^2, inv 2 N^2 per slice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vsmartmom.jl_b200"))
import torch
from vsmom_b200 import _lib, batched
ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
for dt in (torch.float64, torch.float32):
    for N, batch in ((45, 20000), (15, 100000), (100, 4000)):
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.rand((batch, N, N), dtype=dt, device="cuda", generator=g) * 0.1 + torch.eye(N, dtype=dt, device="cuda")
        B = torch.rand((batch, N, N), dtype=dt, device="cuda", generator=g)
        X = torch.empty_like(A)
        esz = A.element_size()
        def timeit(fn, reps=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        tm = timeit(lambda: batched.batched_mul(ctx, A, B))
        ti = timeit(lambda: batched.batch_inv_(ctx, X, A.clone()))
        tc = timeit(lambda: A.clone())
        tt = timeit(lambda: torch.bmm(A, B))
        print(f"{str(dt):14s} N={N:3d} batch={batch:6d}  mul {tm:7.3f} ms = {3*N*N*esz*batch/tm/1e6:7.0f} GB/s ({2*N**3*batch/tm/1e9:6.2f} TF)   "
              f"torch.bmm {tt:7.3f} ms   inv(+clone) {ti:7.3f} ms (clone {tc:.3f}) = {2*N*N*esz*batch/max(ti-tc,1e-6)/1e6:7.0f} GB/s")

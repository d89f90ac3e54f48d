"""Test infrastructure (developer check, ~2 min, CPU only): is the truncated Neumann evaluation of
(I - A)^-1 that the fp64 kernels use for small ||A||_F as accurate as the LAPACK inverse?

Runs the oracle three times on a thick, bright variant of the N=45 / 34-layer scene:
  (a) float64 with numpy/LAPACK inv (the oracle as committed),
  (b) float64 with the kernels' route selection (K = 2 / 4 / 8 Neumann product form, else LAPACK),
  (c) 80-bit long double with an unpivoted Gauss-Jordan (the judge between (a) and (b)),
and prints max|x - c| / max|c| for (a) and (b).  Measured (scale 8):  R_SFI 6.4e-12 vs 8.5e-12,
T_SFI 6.7e-12 vs 1.1e-11, hdr 9.6e-12 vs 1.2e-11 -- both at the rounding floor of the problem.

    python oracle/check_inverse_routes_longdouble.py 8.0
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in ("oracle", "vsmartmom.jl_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, _p))
import vsmom_oracle as o
import vsmom_b200 as vb
from helpers import to_oracle, relerr
LD = np.longdouble
def gj_inv(X):
    X = np.array(X, dtype=X.dtype, copy=True)
    n = X.shape[-1]
    for k in range(n):
        p = X[..., k, k].copy()
        ip = 1 / p
        row = X[..., k, :] * ip[..., None]
        col = X[..., :, k].copy()
        X -= col[..., :, None] * row[..., None, :]
        X[..., k, :] = row
        X[..., :, k] = -col * ip[..., None]
        X[..., k, k] = ip
    return X
def neumann_inv(X):
    n = X.shape[-1]; I = np.eye(n, dtype=X.dtype)
    A = I - X
    nu2 = (A*A).sum(axis=(-1,-2))
    out = np.empty_like(X)
    for b in np.ndindex(nu2.shape):
        a = A[b]; v = nu2[b]
        if v <= 1e-20: out[b] = I + a
        elif v <= 1e-10:
            a2 = a@a; out[b] = (I + a) + (a2 + a@a2)
        elif v <= 9e-6:
            a2 = a@a; P = (I + a) + (a2 + a@a2); a4 = a2@a2; out[b] = P + P@a4
        else: out[b] = orig(X[b])
    return out
orig = np.linalg.inv
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
def model(ft):
    m = vb.synthetic_o2a_model(nSpec=12, aerosol=True, albedo=0.8, float_type=ft)
    m.tau_rayl = m.tau_rayl*scale; m.tau_aer = m.tau_aer*scale; m.tau_abs = m.tau_abs*np.logspace(-5.0,-0.5,12)[:,None]
    return m
sel = slice(0, 12, 3)
t=time.time()
o.np.linalg.inv = orig;        ref64 = o.rt_run(to_oracle(model(np.float64)), sel)
o.np.linalg.inv = neumann_inv; neu64 = o.rt_run(to_oracle(model(np.float64)), sel)
print('f64 done', time.time()-t)
o.np.linalg.inv = gj_inv
try:
    refld = o.rt_run(to_oracle(model(LD)), sel)
    print('ld done', time.time()-t, refld['R_SFI'].dtype)
    for k in ('R_SFI','T_SFI','hdr'):
        print(k, 'lapack64-ld %.2e  neumann64-ld %.2e  neumann64-lapack64 %.2e' % (relerr(ref64[k], refld[k].astype(np.float64)), relerr(neu64[k], refld[k].astype(np.float64)), relerr(neu64[k], ref64[k])))
except Exception as e:
    import traceback; traceback.print_exc()

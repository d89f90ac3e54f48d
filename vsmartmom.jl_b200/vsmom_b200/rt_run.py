"""``rt_run`` for the B200 path -- host-side mirror of
``rt_run(::noRS, model, iBand)`` (src/CoreRT/rt_run.jl:238-539): same inputs,
same outputs ``(R_SFI, T_SFI, hdr, bhr_uw, bhr_dw)``; the Fourier/layer loop runs
in ``vsm_rt_column_*`` on the GPU.  PyTorch is used only for device memory and
streams."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .model import RTModel, _prt_pm, compute_Z_moments, layer_control, vza_tables


def _tdtype(FT):
    return torch.float64 if FT == np.float64 else torch.float32


@dataclass
class PreparedColumn:
    """Scene-constant inputs (everything except the spectral arrays): the Z-moment
    tables, quadrature and full-band control flow.  Built once per model."""
    model: RTModel
    Zpp: torch.Tensor
    Zmp: torch.Tensor
    ctl: object
    vza_stream: np.ndarray
    vza_weight: np.ndarray
    surf_rho: Optional[torch.Tensor] = None      # (n_moments, N, N), memory order == Julia (N,N,n_moments)
    ss_correction: Optional[np.ndarray] = None   # (pol_n, nVZA) float64, memory order == (nVZA, nStokes) column-major


def z_tables(model: RTModel) -> tuple[np.ndarray, np.ndarray]:
    """(N,N,1+nAer,n_moments) column-major Z⁺⁺ / Z⁻⁺ tables in the model's FT."""
    mu = model.quad_points.qp_mu.astype(np.float64)
    comps = [model.greek_rayleigh] + [a.greek_coefs for a in model.aerosol_optics]
    N = model.quad_points.qp_muN.size
    nM = model.m_max + 1
    Zpp = np.zeros((nM, len(comps), N, N), model.float_type)
    Zmp = np.zeros_like(Zpp)
    lmax = max(len(g.beta) for g in comps)
    for m in range(nM):
        prt = _prt_pm(mu, lmax, m)                          # shared by every scatterer
        for c, g in enumerate(comps):
            zp, zm = compute_Z_moments(model.pol_n, mu, g, m, prt)
            # memory order (m, c, j, i) == Julia (i, j, c, m)
            Zpp[m, c] = zp.T
            Zmp[m, c] = zm.T
    return Zpp, Zmp


def prepare(model: RTModel, device: torch.device | str = "cuda:0") -> PreparedColumn:
    Zpp, Zmp = z_tables(model)
    stream, w = vza_tables(model)
    rho = None
    if model.surface_rho is not None:
        N = model.quad_points.qp_muN.size
        sr = np.asarray(model.surface_rho)
        if sr.shape != (model.m_max + 1, N, N):
            raise ValueError(f"surface_rho must have shape (m_max+1, N, N) = {(model.m_max + 1, N, N)}, got {sr.shape}")
        # [m, j, i] <- rho_m[i, j]: column-major (N, N, n_moments) for the library
        rho = torch.from_numpy(np.ascontiguousarray(sr.transpose(0, 2, 1).astype(model.float_type))).to(device)
    ss = None
    if model.ss_correction is not None:
        sc = np.asarray(model.ss_correction, np.float64)
        if sc.shape != (len(model.vza), model.pol_n):
            raise ValueError(f"ss_correction must have shape (nVZA, pol_n) = {(len(model.vza), model.pol_n)}, got {sc.shape}")
        ss = np.ascontiguousarray(sc.T)
    return PreparedColumn(model, torch.from_numpy(Zpp).to(device), torch.from_numpy(Zmp).to(device),
                          layer_control(model), stream, np.ascontiguousarray(w, dtype=np.float64), rho, ss)


def _f64p(a: np.ndarray):
    return a.ctypes.data_as(_lib.f64p)


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(_lib.i32p)


def _column_desc(prep: PreparedColumn, tau_rayl: torch.Tensor, tau_abs: torch.Tensor, F0: torch.Tensor):
    """The vsm_column_desc of one call plus the host arrays it points to (which must stay alive)."""
    model = prep.model
    q = model.quad_points
    Nz, nS = tau_rayl.shape
    nV, n = len(model.vza), model.pol_n
    mu = np.ascontiguousarray(q.qp_muN, dtype=np.float64)
    wt = np.ascontiguousarray(q.wt_muN, dtype=np.float64)
    nAer = len(model.aerosol_optics)
    om = np.array([a.omega for a in model.aerosol_optics] or [0.0], np.float64)
    ft = np.array([a.ft for a in model.aerosol_optics] or [0.0], np.float64)
    taer = np.ascontiguousarray(model.tau_aer.T if nAer else np.zeros((1, 1)), dtype=np.float64)  # (Nz, nAer) mem == Julia (nAer,Nz)
    ctl = prep.ctl
    nd = np.ascontiguousarray(ctl.ndoubl, np.int32); sc = np.ascontiguousarray(ctl.scatter, np.int32)
    itf = np.ascontiguousarray(ctl.iface, np.int32); rz = np.ascontiguousarray(ctl.rayl_allzero, np.int32)
    vs = np.ascontiguousarray(prep.vza_stream, np.int32)
    d = _lib.ColumnDesc()
    d.struct_size = C.sizeof(_lib.ColumnDesc)
    d.N, d.nStokes, d.nSpec, d.Nz, d.nAer, d.nVZA = q.qp_muN.size, n, nS, Nz, nAer, nV
    d.n_moments, d.imu0, d.surface_kind = model.m_max + 1, q.imu0, (1 if prep.surf_rho is not None else 0)
    d.surf_rho = prep.surf_rho.data_ptr() if prep.surf_rho is not None else None
    d.ss_correction = _f64p(prep.ss_correction) if prep.ss_correction is not None else None
    d.mu0, d.albedo, d.varpi_cabannes = float(q.mu0), float(model.albedo), float(model.varpi_cabannes)
    d.mu, d.wt, d.aer_omega, d.aer_ft, d.tau_aer = _f64p(mu), _f64p(wt), _f64p(om), _f64p(ft), _f64p(taer)
    d.vza_weight = _f64p(prep.vza_weight)
    d.ndoubl, d.scatter, d.iface, d.rayl_allzero, d.vza_stream = _i32p(nd), _i32p(sc), _i32p(itf), _i32p(rz), _i32p(vs)
    d.tau_rayl, d.tau_abs, d.F0 = tau_rayl.data_ptr(), tau_abs.data_ptr(), F0.data_ptr()
    d.Zpp, d.Zmp = prep.Zpp.data_ptr(), prep.Zmp.data_ptr()
    return d, (mu, wt, om, ft, taer, nd, sc, itf, rz, vs)


def rt_column_device(ctx: _lib.Context, prep: PreparedColumn, tau_rayl: torch.Tensor, tau_abs: torch.Tensor,
                     F0: torch.Tensor, want_hdr: bool = True, streams: Optional[dict] = None,
                     thermal_B: Optional[torch.Tensor] = None, sif0: Optional[torch.Tensor] = None):
    """Device-resident entry: ``tau_rayl``/``tau_abs`` are (Nz, nS) torch CUDA tensors
    (memory order == Julia (nSpec, Nz)); ``F0`` is (nS, pol_n).  Returns device
    tensors R_SFI, T_SFI, hdr of shape (nS, pol_n, nVZA) (memory order == Julia
    (nVZA, pol_n, nSpec)) and bhr_uw, bhr_dw of shape (nS, pol_n).
    ``streams``: pass an empty dict to also get the per-moment composite at all streams
    (``rt_run_streams``): it is filled with device tensors ``Rmp``, ``Tpp`` of shape
    (n_moments, nS, N, N) [memory order == Julia (N,N,nSpec) per moment] and ``Jm``, ``Jp``
    of shape (n_moments, nS, N)."""
    model = prep.model
    FT = model.float_type
    dt = _tdtype(FT)
    q = model.quad_points
    Nz, nS = tau_rayl.shape
    assert tau_abs.shape == (Nz, nS) and tau_rayl.dtype == dt and tau_rayl.is_contiguous() and tau_abs.is_contiguous()
    dev = tau_rayl.device
    nV, n = len(model.vza), model.pol_n
    R = torch.empty((nS, n, nV), dtype=dt, device=dev)
    T = torch.empty_like(R)
    H = torch.empty_like(R) if want_hdr else None
    uw = torch.empty((nS, n), dtype=dt, device=dev) if want_hdr else None
    dw = torch.empty((nS, n), dtype=dt, device=dev) if want_hdr else None
    d, _keep = _column_desc(prep, tau_rayl, tau_abs, F0)
    if thermal_B is not None:           # (nS, Nz) tensor, memory == Julia (Nz, nSpec)
        assert thermal_B.shape == (nS, Nz) and thermal_B.dtype == dt and thermal_B.is_contiguous()
        d.thermal_B = thermal_B.data_ptr()
    if sif0 is not None:                # (nS, pol_n) tensor, memory == Julia (nStokes, nSpec)
        assert sif0.shape == (nS, n) and sif0.dtype == dt and sif0.is_contiguous()
        d.sif0 = sif0.data_ptr()
    hp = (H.data_ptr() if want_hdr else None, uw.data_ptr() if want_hdr else None, dw.data_ptr() if want_hdr else None)
    if streams is not None:
        N, nM = q.qp_muN.size, model.m_max + 1
        streams["Rmp"] = torch.empty((nM, nS, N, N), dtype=dt, device=dev)
        streams["Tpp"] = torch.empty((nM, nS, N, N), dtype=dt, device=dev)
        streams["Jm"] = torch.empty((nM, nS, N), dtype=dt, device=dev)
        streams["Jp"] = torch.empty((nM, nS, N), dtype=dt, device=dev)
        fn = _lib.lib.vsm_rt_column_streams_f64 if FT == np.float64 else _lib.lib.vsm_rt_column_streams_f32
        ctx.check(fn(ctx.handle, C.byref(d), R.data_ptr(), T.data_ptr(), *hp, streams["Rmp"].data_ptr(),
                     streams["Tpp"].data_ptr(), streams["Jm"].data_ptr(), streams["Jp"].data_ptr()))
    else:
        fn = _lib.lib.vsm_rt_column_f64 if FT == np.float64 else _lib.lib.vsm_rt_column_f32
        ctx.check(fn(ctx.handle, C.byref(d), R.data_ptr(), T.data_ptr(), *hp))
    return R, T, H, uw, dw


def rt_run(model: RTModel, ctx: Optional[_lib.Context] = None, prep: Optional[PreparedColumn] = None,
           device: str = "cuda:0", sel: slice = slice(None), F0: Optional[np.ndarray] = None,
           return_streams: bool = False):
    """Host-buffer entry, the call a user makes: NumPy in, NumPy out, like the
    reference's ``rt_run(model)``.  Returns a dict with ``R_SFI``, ``T_SFI``,
    ``hdr`` of shape (nVZA, pol_n, nSpec) and ``bhr_uw``, ``bhr_dw`` (nSpec,).
    ``sel`` restricts the spectral axis (a shard); control flow always comes from
    the full band.  ``return_streams`` adds ``streams``: per Fourier moment the composite
    ``Rmp``, ``Tpp`` (n_moments, nSpec, N, N) [row i, column j] and ``Jm``, ``Jp``
    (n_moments, nSpec, N) at all quadrature streams (the reference's ``rt_run_streams``)."""
    if not torch.cuda.is_available():
        raise RuntimeError("vsmom_b200.rt_run needs a CUDA device (B200); there is no CPU fallback")
    own = ctx is None
    if own:
        ctx = _lib.Context(torch.device(device).index or 0)
    if prep is None:
        prep = prepare(model, device)
    FT = model.float_type
    dt = _tdtype(FT)
    tr = np.ascontiguousarray(model.tau_rayl[sel].T)      # (Nz, nS)
    ta = np.ascontiguousarray(model.tau_abs[sel].T)
    nS = tr.shape[1]
    if F0 is None:
        f0 = np.zeros((nS, model.pol_n), FT)
        f0[:, 0] = 1
    else:
        f0 = np.ascontiguousarray(F0[:, sel].T.astype(FT))
    with torch.cuda.device(device):
        d_tr = torch.from_numpy(tr).pin_memory().to(device, non_blocking=True)
        d_ta = torch.from_numpy(ta).pin_memory().to(device, non_blocking=True)
        d_f0 = torch.from_numpy(f0).to(device)
        st = {} if return_streams else None
        d_th = None if model.thermal_B is None else torch.from_numpy(np.ascontiguousarray(model.thermal_B[sel].astype(FT))).to(device)
        d_sif = None if model.sif0 is None else torch.from_numpy(np.ascontiguousarray(model.sif0[sel].astype(FT))).to(device)
        R, T, H, uw, dw = rt_column_device(ctx, prep, d_tr, d_ta, d_f0, streams=st, thermal_B=d_th, sif0=d_sif)
        ctx.synchronize()
        out = dict(R_SFI=R.cpu().numpy().transpose(2, 1, 0), T_SFI=T.cpu().numpy().transpose(2, 1, 0),
                   hdr=H.cpu().numpy().transpose(2, 1, 0), bhr_uw=uw.cpu().numpy()[:, 0].copy(),
                   bhr_dw=dw.cpu().numpy()[:, 0].copy())
        if return_streams:      # device memory is (m, s, j, i): hand out [m, s, i, j]
            out["streams"] = dict(Rmp=st["Rmp"].cpu().numpy().transpose(0, 1, 3, 2),
                                  Tpp=st["Tpp"].cpu().numpy().transpose(0, 1, 3, 2),
                                  Jm=st["Jm"].cpu().numpy(), Jp=st["Jp"].cpu().numpy())
    if own:
        ctx.close()
    return out


def rt_run_ss(model: RTModel, ctx: Optional[_lib.Context] = None, device: str = "cuda:0"):
    """Single-scattering driver, the reference's ``rt_run_ss(model)`` (rt_run.jl:547-755, noRS): returns ``R_SFI``,
    ``T_SFI`` (nVZA, pol_n, nSpec) and the hemispheric ``hem_R``, ``hem_T`` (nSpec,)."""
    if not torch.cuda.is_available():
        raise RuntimeError("vsmom_b200.rt_run_ss needs a CUDA device (B200); there is no CPU fallback")
    own = ctx is None
    if own:
        ctx = _lib.Context(torch.device(device).index or 0)
    prep = prepare(model, device)
    FT = model.float_type
    dt = _tdtype(FT)
    nS, n, nV = model.nSpec, model.pol_n, len(model.vza)
    with torch.cuda.device(device):
        tr = torch.from_numpy(np.ascontiguousarray(model.tau_rayl.T)).to(device)
        ta = torch.from_numpy(np.ascontiguousarray(model.tau_abs.T)).to(device)
        f0 = np.zeros((nS, n), FT); f0[:, 0] = 1
        f0_d = torch.from_numpy(f0).to(device)            # (kept alive: the descriptor holds its raw pointer)
        d, _keep = _column_desc(prep, tr, ta, f0_d)
        R = torch.empty((nS, n, nV), dtype=dt, device=device); T = torch.empty_like(R)
        hR = torch.empty((nS,), dtype=dt, device=device); hT = torch.empty_like(hR)
        fn = _lib.lib.vsm_rt_column_ss_f64 if FT == np.float64 else _lib.lib.vsm_rt_column_ss_f32
        ctx.check(fn(ctx.handle, C.byref(d), R.data_ptr(), T.data_ptr(), hR.data_ptr(), hT.data_ptr()))
        ctx.synchronize()
        out = dict(R_SFI=R.cpu().numpy().transpose(2, 1, 0), T_SFI=T.cpu().numpy().transpose(2, 1, 0),
                   hem_R=hR.cpu().numpy(), hem_T=hT.cpu().numpy())
    if own:
        ctx.close()
    return out


def rt_run_sharded(model: RTModel, ctx: Optional[_lib.Context] = None, device: Optional[str] = None):
    """``rt_run`` with the spectral axis sharded over the ranks of the default
    torch.distributed process group (NCCL, one process per GPU).  Every rank passes
    the same full-band model and gets the full-band result."""
    import torch.distributed as dist
    from .sharding import gather_spectral, shard_bounds
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    device = device or f"cuda:{torch.cuda.current_device()}"
    own = ctx is None
    if own:
        ctx = _lib.Context(torch.device(device).index or 0)
    prep = prepare(model, device)                      # full-band control flow
    lo, hi = shard_bounds(model.nSpec, world)[rank]
    FT = model.float_type
    nS = hi - lo
    dt = _tdtype(FT)
    nV, n = len(model.vza), model.pol_n
    if nS > 0:
        d_tr = torch.from_numpy(np.ascontiguousarray(model.tau_rayl[lo:hi].T)).to(device)
        d_ta = torch.from_numpy(np.ascontiguousarray(model.tau_abs[lo:hi].T)).to(device)
        f0 = np.zeros((nS, n), FT); f0[:, 0] = 1
        d_th = None if model.thermal_B is None else torch.from_numpy(np.ascontiguousarray(model.thermal_B[lo:hi].astype(FT))).to(device)
        d_sif = None if model.sif0 is None else torch.from_numpy(np.ascontiguousarray(model.sif0[lo:hi].astype(FT))).to(device)
        d_f0 = torch.from_numpy(f0).to(device)             # inputs stay alive until the gather has synchronised
        R, T, H, uw, dw = rt_column_device(ctx, prep, d_tr, d_ta, d_f0, thermal_B=d_th, sif0=d_sif)
        ctx.synchronize()
    else:
        R = torch.zeros((0, n, nV), dtype=dt, device=device); T = R.clone(); H = R.clone()
        uw = torch.zeros((0, n), dtype=dt, device=device); dw = uw.clone()
    packed = torch.cat([R.reshape(nS, -1), T.reshape(nS, -1), H.reshape(nS, -1), uw, dw], dim=1)
    full = gather_spectral(packed, model.nSpec, world, rank).cpu().numpy()
    k = n * nV
    out = dict(R_SFI=full[:, :k].reshape(-1, n, nV).transpose(2, 1, 0),
               T_SFI=full[:, k:2 * k].reshape(-1, n, nV).transpose(2, 1, 0),
               hdr=full[:, 2 * k:3 * k].reshape(-1, n, nV).transpose(2, 1, 0),
               bhr_uw=full[:, 3 * k].copy(), bhr_dw=full[:, 3 * k + n].copy())
    if own:
        ctx.close()
    return out


def rt_run_multi(model: RTModel, mctx: "_lib.MultiContext"):
    """``rt_run`` on several GPUs driven from THIS process (``vsm_rt_column_sharded_*``): host arrays in, host arrays
    out; the library shards the spectral axis over the devices of ``mctx`` and gathers with one NCCL all-gather."""
    FT = model.float_type
    Zpp, Zmp = z_tables(model)
    stream, w = vza_tables(model)
    prep = PreparedColumn(model, torch.from_numpy(Zpp), torch.from_numpy(Zmp), layer_control(model), stream,
                          np.ascontiguousarray(w, dtype=np.float64), None,
                          None if model.ss_correction is None else np.ascontiguousarray(np.asarray(model.ss_correction, np.float64).T))
    if model.surface_rho is not None:
        prep.surf_rho = torch.from_numpy(np.ascontiguousarray(np.asarray(model.surface_rho).transpose(0, 2, 1).astype(FT)))
    nS, n, nV = model.nSpec, model.pol_n, len(model.vza)
    tr = torch.from_numpy(np.ascontiguousarray(model.tau_rayl.T))       # (Nz, nSpec) mem == Julia (nSpec, Nz), HOST
    ta = torch.from_numpy(np.ascontiguousarray(model.tau_abs.T))
    f0 = np.zeros((nS, n), FT); f0[:, 0] = 1
    F0 = torch.from_numpy(f0)
    d, _keep = _column_desc(prep, tr, ta, F0)
    th = None if model.thermal_B is None else np.ascontiguousarray(model.thermal_B.astype(FT))       # (nSpec, Nz), HOST
    sf = None if model.sif0 is None else np.ascontiguousarray(model.sif0.astype(FT))
    if th is not None:
        d.thermal_B = th.ctypes.data
    if sf is not None:
        d.sif0 = sf.ctypes.data
    R = np.empty((nS, n, nV), FT); T = np.empty_like(R); H = np.empty_like(R)
    uw = np.empty((nS, n), FT); dw = np.empty_like(uw)
    fn = _lib.lib.vsm_rt_column_sharded_f64 if FT == np.float64 else _lib.lib.vsm_rt_column_sharded_f32
    mctx.check(fn(mctx.handle, C.byref(d), R.ctypes.data, T.ctypes.data, H.ctypes.data, uw.ctypes.data, dw.ctypes.data))
    return dict(R_SFI=R.transpose(2, 1, 0), T_SFI=T.transpose(2, 1, 0), hdr=H.transpose(2, 1, 0),
                bhr_uw=uw[:, 0].copy(), bhr_dw=dw[:, 0].copy())


def rt_run_lin(model: RTModel, dtau_rayl: Optional[np.ndarray] = None, dtau_abs: Optional[np.ndarray] = None,
               dtau_aer: Optional[np.ndarray] = None, dalbedo: Optional[np.ndarray] = None,
               ctx: Optional[_lib.Context] = None, device: str = "cuda:0"):
    """LinMode: forward ``R_SFI``, ``T_SFI`` (nVZA, pol_n, nSpec) and the analytic Jacobians ``dR_SFI``, ``dT_SFI``
    (P, nVZA, pol_n, nSpec) -- the reference's ``rt_run(model, lin_model, ...)`` (rt_run_lin.jl:102-326, outputs
    :185-190) for P parameters given by the derivatives of the column inputs: ``dtau_rayl``, ``dtau_abs``
    (P, nSpec, Nz), ``dtau_aer`` (P, nAer, Nz), ``dalbedo`` (P,); any may be None (= 0).  fp64."""
    if not torch.cuda.is_available():
        raise RuntimeError("vsmom_b200.rt_run_lin needs a CUDA device (B200); there is no CPU fallback")
    if model.float_type != np.float64:
        raise ValueError("rt_run_lin is fp64 only")
    given = [x for x in (dtau_rayl, dtau_abs, dtau_aer, dalbedo) if x is not None]
    if not given:
        raise ValueError("rt_run_lin: no parameter given")
    P = int(np.asarray(given[0]).shape[0])
    own = ctx is None
    if own:
        ctx = _lib.Context(torch.device(device).index or 0)
    prep = prepare(model, device)
    nS, Nz, n, nV = model.nSpec, model.Nz, model.pol_n, len(model.vza)
    dev = torch.device(device)

    def spectral(x):        # (P, nSpec, Nz) -> memory (P, Nz, nSpec) == Julia (nSpec, Nz, P)
        if x is None:
            return None
        x = np.asarray(x, np.float64)
        assert x.shape == (P, nS, Nz), x.shape
        return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1))).to(dev)
    d_tr, d_ta = spectral(dtau_rayl), spectral(dtau_abs)
    tr = torch.from_numpy(np.ascontiguousarray(model.tau_rayl.T)).to(dev)
    ta = torch.from_numpy(np.ascontiguousarray(model.tau_abs.T)).to(dev)
    f0 = np.zeros((nS, n)); f0[:, 0] = 1
    F0 = torch.from_numpy(f0).to(dev)
    d, _keep = _column_desc(prep, tr, ta, F0)
    ld = _lib.LinDesc()
    ld.struct_size, ld.n_params = C.sizeof(_lib.LinDesc), P
    aer = None
    if dtau_aer is not None and len(model.aerosol_optics):
        aer = np.ascontiguousarray(np.asarray(dtau_aer, np.float64).transpose(0, 2, 1))     # (P, Nz, nAer) mem == (nAer,Nz,P)
        assert aer.shape == (P, Nz, len(model.aerosol_optics))
        ld.dtau_aer = _f64p(aer)
    alb = None
    if dalbedo is not None:
        alb = np.ascontiguousarray(np.asarray(dalbedo, np.float64))
        assert alb.shape == (P,)
        ld.dalbedo = _f64p(alb)
    ld.dtau_rayl = d_tr.data_ptr() if d_tr is not None else None
    ld.dtau_abs = d_ta.data_ptr() if d_ta is not None else None
    R = torch.empty((nS, n, nV), dtype=torch.float64, device=dev)
    T = torch.empty_like(R)
    dR = torch.empty((P, nS, n, nV), dtype=torch.float64, device=dev)
    dT = torch.empty_like(dR)
    ctx.check(_lib.lib.vsm_rt_column_lin_f64(ctx.handle, C.byref(d), C.byref(ld), R.data_ptr(), T.data_ptr(),
                                             dR.data_ptr(), dT.data_ptr()))
    ctx.synchronize()
    out = dict(R_SFI=R.cpu().numpy().transpose(2, 1, 0), T_SFI=T.cpu().numpy().transpose(2, 1, 0),
               dR_SFI=dR.cpu().numpy().transpose(0, 3, 2, 1), dT_SFI=dT.cpu().numpy().transpose(0, 3, 2, 1))
    if own:
        ctx.close()
    return out


@dataclass
class StreamRTResult:
    """Mirror of the reference's ``StreamRTResult`` (rt_run.jl:60-131): per-Fourier-moment result
    at all internal quadrature streams."""
    qp_mu: np.ndarray
    imu0: int                 # 1-based, as in the reference
    mu0: float
    pol_n: int
    weight: np.ndarray        # 0.5/pi for m = 0, 1/pi otherwise
    Rmp_per_m: np.ndarray     # (n_moments, nSpec, N, N)
    Tpp_per_m: np.ndarray
    Jm_per_m: np.ndarray      # (n_moments, nSpec, N)
    Jp_per_m: np.ndarray
    tau_total: np.ndarray     # (nSpec, Nz) = tau_rayl + tau_abs (aerosol not included, as in the reference)
    tau_rayl: np.ndarray
    tau_abs: np.ndarray


def rt_run_streams(model: RTModel, ctx: Optional[_lib.Context] = None, device: str = "cuda:0") -> StreamRTResult:
    """``rt_run_streams(model)`` (rt_run.jl:133-193)."""
    out = rt_run(model, ctx, device=device, return_streams=True)
    s = out["streams"]
    q = model.quad_points
    nM = model.m_max + 1
    w = np.array([0.5 / np.pi] + [1.0 / np.pi] * (nM - 1), model.float_type)
    return StreamRTResult(np.asarray(q.qp_mu), q.imu0, float(q.mu0), model.pol_n, w, s["Rmp"], s["Tpp"], s["Jm"], s["Jp"],
                          np.asarray(model.tau_rayl + model.tau_abs), np.asarray(model.tau_rayl), np.asarray(model.tau_abs))

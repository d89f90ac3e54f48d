"""Rotational Raman scattering (RRS) on the B200 path -- host-side mirror of the reference's
``rt_run(RS_type::RRS, model, iBand)`` (src/CoreRT/rt_run.jl:238-539, rt_kernel!(::RRS) CoreKernel/rt_kernel.jl:352-391):
same inputs (the model plus the fields of ``InelasticScattering.RRS`` the hot path reads), same outputs
``(R_SFI, T_SFI, ieR_SFI, ieT_SFI)``; the Fourier / layer / Raman-shift loops run in ``vsm_rt_column_rrs_f64``.

The Raman line physics (``getRamanSSProp!``, raman_atmo_prop.jl:80-100: N2/O2 constants -> ϖ_λ₁λ₀, i_λ₁λ₀) is not
part of the hot path; ``synthetic_rrs`` builds a line table of the same shape for the synthetic band."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .model import (AerosolOptics, GreekCoefs, RTModel, RTNumericalParameters, compute_Z_moments, get_greek_rayleigh,
                    rt_set_streams, bodhaine_tau_rayleigh, _synthetic_absorption, POL_N)
from .rt_run import _column_desc, _f64p, _i32p, prepare


@dataclass
class RRS:
    """The fields of ``InelasticScattering.RRS`` (src/Inelastic/types.jl) read by the CoreRT loop."""
    i_shift: np.ndarray        # i_λ₁λ₀ [nRaman]: n₀ = n₁ + i_shift[Δn]
    varpi_raman: np.ndarray    # ϖ_λ₁λ₀ [nRaman] (see normalize_raman_weights)
    greek_raman: GreekCoefs


def get_greek_raman(depol: float) -> GreekCoefs:
    """``get_greek_raman`` (inelastic_helper.jl:862-880) for the rotational-Raman depolarisation ratio."""
    dpl_p = (1 - depol) / (1 + depol / 2)
    dpl_r = (1 - 2 * depol) / (1 - depol)
    z = np.zeros(3)
    return GreekCoefs(np.array([0.0, 0.0, 3 * dpl_p]), np.array([1.0, 0.0, 0.5 * dpl_p]),
                      np.array([0.0, 0.0, dpl_p * math.sqrt(1.5)]), np.array([0.0, dpl_p * dpl_r * 1.5, 0.0]), z, z.copy())


def normalize_raman_weights(varpi_raman, varpi_cabannes: float) -> np.ndarray:
    """``normalize_raman_weights!`` (Inelastic/types.jl:625-629): Σ ϖ_λ₁λ₀ = 1 - ϖ_Cabannes."""
    w = np.asarray(varpi_raman, np.float64)
    return w * ((1.0 - varpi_cabannes) / np.sum(w))


def raman_z_tables(model: RTModel, rs: RRS) -> Tuple[np.ndarray, np.ndarray]:
    """``computeRamanZλ!`` (inelastic_helper.jl:917-924) for every Fourier moment: memory (m, j, i) == Julia (N,N,n_moments)."""
    mu = model.quad_points.qp_mu.astype(np.float64)
    N = model.quad_points.qp_muN.size
    nM = model.m_max + 1
    Zpp = np.zeros((nM, N, N))
    Zmp = np.zeros_like(Zpp)
    for m in range(nM):
        zp, zm = compute_Z_moments(model.pol_n, mu, rs.greek_raman, m)
        Zpp[m] = zp.T
        Zmp[m] = zm.T
    return Zpp, Zmp


def rrs_column_device(ctx: _lib.Context, prep, rs: RRS, tau_rayl: torch.Tensor, tau_abs: torch.Tensor, F0: torch.Tensor,
                      ZppR: torch.Tensor, ZmpR: torch.Tensor, n1_begin: int, n1_count: int):
    """Device-resident entry.  ``tau_rayl``/``tau_abs`` (Nz, nBand) and ``F0`` (nBand, pol_n) cover the WHOLE band;
    the outputs (n1_count, pol_n, nVZA) [memory == Julia (nVZA, pol_n, n1_count)] cover the owner range."""
    model = prep.model
    if model.float_type != np.float64:
        raise ValueError("the Raman path is fp64 only")
    nV, n = len(model.vza), model.pol_n
    dev = tau_rayl.device
    d, keep = _column_desc(prep, tau_rayl, tau_abs, F0)
    sh = np.ascontiguousarray(rs.i_shift, np.int32)
    wr = np.ascontiguousarray(rs.varpi_raman, np.float64)
    rd = _lib.RrsDesc()
    rd.struct_size, rd.n_raman, rd.n1_begin, rd.n1_count = C.sizeof(_lib.RrsDesc), int(sh.size), int(n1_begin), int(n1_count)
    rd.i_shift, rd.varpi_raman = _i32p(sh), _f64p(wr)
    rd.Zpp_raman, rd.Zmp_raman = ZppR.data_ptr(), ZmpR.data_ptr()
    outs = [torch.empty((n1_count, n, nV), dtype=torch.float64, device=dev) for _ in range(4)]
    ctx.check(_lib.lib.vsm_rt_column_rrs_f64(ctx.handle, C.byref(d), C.byref(rd), *[o.data_ptr() for o in outs]))
    return outs


def rt_run_rrs(model: RTModel, rs: RRS, ctx: Optional[_lib.Context] = None, device: str = "cuda:0",
               owner: Optional[Tuple[int, int]] = None, F0: Optional[np.ndarray] = None):
    """Host-buffer entry, the call a user makes.  Returns ``R_SFI``, ``T_SFI``, ``ieR_SFI``, ``ieT_SFI`` of shape
    (nVZA, pol_n, n) for the owner range ``owner = (lo, hi)`` of band indices (default: the whole band).
    ``model.greek_rayleigh`` must be the Cabannes set and ``model.varpi_cabannes`` ϖ_Cabannes."""
    if not torch.cuda.is_available():
        raise RuntimeError("vsmom_b200.rt_run_rrs needs a CUDA device (B200); there is no CPU fallback")
    own = ctx is None
    if own:
        ctx = _lib.Context(torch.device(device).index or 0)
    prep = prepare(model, device)
    lo, hi = owner if owner is not None else (0, model.nSpec)
    n = model.pol_n
    with torch.cuda.device(device):
        tr = torch.from_numpy(np.ascontiguousarray(model.tau_rayl.T)).pin_memory().to(device, non_blocking=True)
        ta = torch.from_numpy(np.ascontiguousarray(model.tau_abs.T)).pin_memory().to(device, non_blocking=True)
        if F0 is None:
            f0 = np.zeros((model.nSpec, n)); f0[:, 0] = 1
        else:
            f0 = np.ascontiguousarray(np.asarray(F0, np.float64).T)
        zp, zm = raman_z_tables(model, rs)
        f0_d, zp_d, zm_d = torch.from_numpy(f0).to(device), torch.from_numpy(zp).to(device), torch.from_numpy(zm).to(device)
        outs = rrs_column_device(ctx, prep, rs, tr, ta, f0_d, zp_d, zm_d, lo, hi - lo)     # inputs stay alive until the sync
        ctx.synchronize()
        res = {k: o.cpu().numpy().transpose(2, 1, 0) for k, o in zip(("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI"), outs)}
    if own:
        ctx.close()
    return res


def rt_run_ss_rrs(model: RTModel, rs: RRS, ctx: Optional[_lib.Context] = None, device: str = "cuda:0"):
    """Single-scattering driver with rotational Raman, the reference's ``rt_run_ss(RS_type::RRS, model, iBand)``
    (rt_run.jl:593-755): ``R_SFI``, ``T_SFI``, ``ieR_SFI``, ``ieT_SFI`` (nVZA, pol_n, nSpec), ``hem_R``, ``hem_T`` (nSpec,)."""
    if not torch.cuda.is_available():
        raise RuntimeError("vsmom_b200.rt_run_ss_rrs needs a CUDA device (B200); there is no CPU fallback")
    own = ctx is None
    if own:
        ctx = _lib.Context(torch.device(device).index or 0)
    prep = prepare(model, device)
    nS, n, nV = model.nSpec, model.pol_n, len(model.vza)
    with torch.cuda.device(device):
        tr = torch.from_numpy(np.ascontiguousarray(model.tau_rayl.T)).to(device)
        ta = torch.from_numpy(np.ascontiguousarray(model.tau_abs.T)).to(device)
        f0 = np.zeros((nS, n)); f0[:, 0] = 1
        f0_d = torch.from_numpy(f0).to(device)            # (kept alive: the descriptor holds its raw pointer)
        d, _keep = _column_desc(prep, tr, ta, f0_d)
        zp, zm = raman_z_tables(model, rs)
        zp_d, zm_d = torch.from_numpy(zp).to(device), torch.from_numpy(zm).to(device)
        sh = np.ascontiguousarray(rs.i_shift, np.int32)
        wr = np.ascontiguousarray(rs.varpi_raman, np.float64)
        rd = _lib.RrsDesc()
        rd.struct_size, rd.n_raman, rd.n1_begin, rd.n1_count = C.sizeof(_lib.RrsDesc), int(sh.size), 0, nS
        rd.i_shift, rd.varpi_raman, rd.Zpp_raman, rd.Zmp_raman = _i32p(sh), _f64p(wr), zp_d.data_ptr(), zm_d.data_ptr()
        outs = [torch.empty((nS, n, nV), dtype=torch.float64, device=device) for _ in range(4)]
        hR = torch.empty((nS,), dtype=torch.float64, device=device); hT = torch.empty_like(hR)
        ctx.check(_lib.lib.vsm_rt_column_ss_rrs_f64(ctx.handle, C.byref(d), C.byref(rd), *[o.data_ptr() for o in outs],
                                                    hR.data_ptr(), hT.data_ptr()))
        ctx.synchronize()
        res = {k: o.cpu().numpy().transpose(2, 1, 0) for k, o in zip(("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI"), outs)}
        res["hem_R"], res["hem_T"] = hR.cpu().numpy(), hT.cpu().numpy()
    if own:
        ctx.close()
    return res


def rrs_unit_weights(nSpec: int, rs: RRS, elastic_cost: float = 2.0) -> np.ndarray:
    """Work per output wavelength: its number of in-band donors (get_n0_n1, inelastic_helper.jl:19-26) plus the elastic column."""
    n1 = np.arange(nSpec)
    w = np.full(nSpec, float(elastic_cost))
    for s in np.asarray(rs.i_shift):
        w += ((n1 + s >= 0) & (n1 + s < nSpec))
    return w


def rrs_shard_bounds(nSpec: int, rs: RRS, world: int):
    """Owner ranges with equal numbers of (n1, dn) units per rank (band-edge wavelengths have fewer in-band donors)."""
    from .sharding import shard_bounds_weighted
    return shard_bounds_weighted(rrs_unit_weights(nSpec, rs), world)


def rt_run_rrs_sharded(model: RTModel, rs: RRS, ctx: Optional[_lib.Context] = None, device: Optional[str] = None):
    """``rt_run_rrs`` with the OWNER range sharded over the ranks of the default process group (one process per
    GPU): each rank recomputes the elastic column of its halo locally (no data-path collective) and ONE all-gather
    of the packed (R | T | ieR | ieT) rows closes the run.  Every rank gets the full band."""
    import torch.distributed as dist
    from .sharding import gather_spectral
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    device = device or f"cuda:{torch.cuda.current_device()}"
    bounds = rrs_shard_bounds(model.nSpec, rs, world)
    lo, hi = bounds[rank]
    n, nV = model.pol_n, len(model.vza)
    if hi > lo:
        r = rt_run_rrs(model, rs, ctx, device, owner=(lo, hi))
        packed = np.concatenate([r[k].transpose(2, 1, 0).reshape(hi - lo, -1) for k in ("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI")], axis=1)
    else:
        packed = np.zeros((0, 4 * n * nV))
    full = gather_spectral(torch.from_numpy(np.ascontiguousarray(packed)).to(device), model.nSpec, world, rank, bounds=bounds).cpu().numpy()
    k = n * nV
    return {name: full[:, i * k:(i + 1) * k].reshape(-1, n, nV).transpose(2, 1, 0)
            for i, name in enumerate(("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI"))}


def rt_run_rrs_multi(model: RTModel, rs: RRS, mctx: "_lib.MultiContext"):
    """``rt_run_rrs`` on several GPUs driven from THIS process (``vsm_rt_column_rrs_sharded_f64``): host arrays in, host
    arrays out; the library gives every device of ``mctx`` one owner range and gathers with one NCCL all-gather."""
    from .rt_run import PreparedColumn, z_tables
    from .model import layer_control, vza_tables
    Zpp, Zmp = z_tables(model)
    stream, w = vza_tables(model)
    prep = PreparedColumn(model, torch.from_numpy(Zpp), torch.from_numpy(Zmp), layer_control(model), stream,
                          np.ascontiguousarray(w, dtype=np.float64), None, None)
    if model.surface_rho is not None:
        prep.surf_rho = torch.from_numpy(np.ascontiguousarray(np.asarray(model.surface_rho).transpose(0, 2, 1).astype(np.float64)))
    nS, n, nV = model.nSpec, model.pol_n, len(model.vza)
    tr = torch.from_numpy(np.ascontiguousarray(model.tau_rayl.T))          # HOST
    ta = torch.from_numpy(np.ascontiguousarray(model.tau_abs.T))
    f0 = np.zeros((nS, n)); f0[:, 0] = 1
    d, _keep = _column_desc(prep, tr, ta, torch.from_numpy(f0))
    zp, zm = raman_z_tables(model, rs)
    sh = np.ascontiguousarray(rs.i_shift, np.int32)
    wr = np.ascontiguousarray(rs.varpi_raman, np.float64)
    rd = _lib.RrsDesc()
    rd.struct_size, rd.n_raman, rd.n1_begin, rd.n1_count = C.sizeof(_lib.RrsDesc), int(sh.size), 0, nS
    rd.i_shift, rd.varpi_raman, rd.Zpp_raman, rd.Zmp_raman = _i32p(sh), _f64p(wr), zp.ctypes.data, zm.ctypes.data
    outs = [np.empty((nS, n, nV)) for _ in range(4)]
    mctx.check(_lib.lib.vsm_rt_column_rrs_sharded_f64(mctx.handle, C.byref(d), C.byref(rd), *[o.ctypes.data for o in outs]))
    return {k: o.transpose(2, 1, 0) for k, o in zip(("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI"), outs)}


# ---------------------------------------------------------------------------
# synthetic config 5 (test/test_forward_raman.jl -> O2Parameters.yaml): IQU, 3 streams + VZA/SZA
# -> N = 15, 5 layers, grid (1e7/775):0.05:(1e7/755) = 6837 points
# ---------------------------------------------------------------------------
def synthetic_rrs(grid_step: float = 0.05, T: float = 250.0, varpi_cabannes: float = 0.96, jmax_n2: int = 24,
                  jmax_o2: int = 29, max_shift_cm: float = 200.0) -> RRS:
    """A rotational-Raman line table with the structure of the reference's (N2: S/O branches ΔJ = ±2 with the 6:3
    nuclear-spin alternation, B = 1.98958 cm⁻¹; O2: odd J only, B = 1.4377 cm⁻¹; Boltzmann populations at T; Placzek-
    Teller factors), shifts rounded onto the spectral grid as ``i_λ₁λ₀`` and weights normalised to 1 - ϖ_Cabannes.
    Shape only -- the cross sections are not those of raman_atmo_prop.jl."""
    kT = 0.6950348 * T
    lines = {}
    for B, jmax, vmr, gam2, par in ((1.98958, jmax_n2, 0.79, 0.51, None), (1.4377, jmax_o2, 0.21, 1.27, 1)):
        Z = sum(((6 if (J % 2 == 0) else 3) if par is None else (1 if J % 2 == par else 0)) * (2 * J + 1) * math.exp(-B * J * (J + 1) / kT)
                for J in range(0, 60))
        for J in range(0, jmax + 1):
            g = (6 if (J % 2 == 0) else 3) if par is None else (1 if J % 2 == par else 0)
            if g == 0:
                continue
            pop = g * (2 * J + 1) * math.exp(-B * J * (J + 1) / kT) / Z
            # Stokes J -> J+2 (photon loses energy) and anti-Stokes J -> J-2
            bS = 3 * (J + 1) * (J + 2) / (2 * (2 * J + 1) * (2 * J + 3))
            dS = B * (4 * J + 6)
            for dnu, b in ((dS, bS),) + (((-B * (4 * J - 2), 3 * J * (J - 1) / (2 * (2 * J + 1) * (2 * J - 1))),) if J >= 2 else ()):
                if abs(dnu) > max_shift_cm:
                    continue
                # the scattered photon at ν₁ came from ν₀ = ν₁ + Δν_Stokes: index offset on an ascending wavenumber grid
                k = int(round(dnu / grid_step))
                if k != 0:
                    lines[k] = lines.get(k, 0.0) + vmr * gam2 * pop * b
    ks = np.array(sorted(lines), np.int32)
    w = np.array([lines[k] for k in ks])
    return RRS(ks, normalize_raman_weights(w, varpi_cabannes), get_greek_raman(0.75 * 0.1))


def synthetic_rrs_model(nSpec: int = 6837, nstreams: int = 3, polarization_type: str = "Stokes_IQU", sza: float = 32.0,
                        vza=(0.0,), vaz=(0.0,), albedo: float = 0.05, Nz: int = 5, varpi_cabannes: float = 0.96,
                        seed: int = 20251017, absorption: float = 5.0) -> RTModel:
    """The config-5 shaped scene: O2-A band on the 0.05 cm⁻¹ grid of O2Parameters.yaml, ``profile_reduction: 5`` layers,
    Cabannes Greek coefficients + ϖ_Cabannes for the elastic Rayleigh part, pseudo-line absorption (the Ring effect
    needs spectral structure to fill in)."""
    pol_n = POL_N[polarization_type]
    q = rt_set_streams(nstreams, sza, list(vza), pol_n, np.float64)
    nu = (1e7 / 775.0) + 0.05 * np.arange(nSpec)
    p_half = np.linspace(0.0, 1013.25, Nz + 1) ** 1.0
    p_half[0] = 0.1
    dp = np.diff(p_half)
    tau_R = bodhaine_tau_rayleigh(1e4 / nu, p_half[-1])
    tau_rayl = tau_R[:, None] * (dp / dp.sum())[None, :]
    span = nu[-1] - nu[0]
    tau_abs = _synthetic_absorption(nu, p_half, seed, nu[0] + 0.06 * span, nu[-1] - 0.06 * span, absorption, nlines=80)
    return RTModel(polarization_type, q, sza, list(vza), list(vaz), tau_rayl, tau_abs, np.zeros((0, Nz)), [],
                   get_greek_rayleigh(0.007), albedo, 2, varpi_cabannes, RTNumericalParameters(), np.float64)

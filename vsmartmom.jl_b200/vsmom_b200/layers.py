"""``AddedLayer`` / ``CompositeLayer`` buffers and ``rt_kernel_`` -- the host-side
mirror of ``rt_kernel!(::noRS, ...)`` (src/CoreRT/CoreKernel/rt_kernel.jl:175-250)
over the C-ABI entry ``vsm_rt_layer_*``.

Tensor convention: torch CUDA tensors whose MEMORY is Julia's column-major
``(N, N, nSpec)`` / ``(N, 1, nSpec)`` array, i.e. torch shape ``(nSpec, N, N)`` with
``A[n, j, i] == A_julia[i, j, n]`` and ``(nSpec, N)`` for the source vectors."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .model import QuadPoints


@dataclass
class AddedLayer:
    """src/CoreRT/types.jl:196-230 (r⁻⁺, t⁺⁺, r⁺⁻, t⁻⁻, j₀⁺, j₀⁻)."""
    rmp: torch.Tensor
    tpp: torch.Tensor
    rpm: torch.Tensor
    tmm: torch.Tensor
    j0p: torch.Tensor
    j0m: torch.Tensor


@dataclass
class CompositeLayer:
    """src/CoreRT/types.jl:155-174 (R⁻⁺, R⁺⁻, T⁺⁺, T⁻⁻, J₀⁺, J₀⁻)."""
    Rmp: torch.Tensor
    Rpm: torch.Tensor
    Tpp: torch.Tensor
    Tmm: torch.Tensor
    J0p: torch.Tensor
    J0m: torch.Tensor


def make_added_layer(N: int, nSpec: int, dtype, device="cuda") -> AddedLayer:
    """make_added_layer (src/CoreRT/tools/rt_helper_functions.jl:130-151)."""
    z3 = lambda: torch.zeros((nSpec, N, N), dtype=dtype, device=device)
    z1 = lambda: torch.zeros((nSpec, N), dtype=dtype, device=device)
    return AddedLayer(z3(), z3(), z3(), z3(), z1(), z1())


def make_composite_layer(N: int, nSpec: int, dtype, device="cuda") -> CompositeLayer:
    """make_composite_layer (rt_helper_functions.jl:259-270)."""
    z3 = lambda: torch.zeros((nSpec, N, N), dtype=dtype, device=device)
    z1 = lambda: torch.zeros((nSpec, N), dtype=dtype, device=device)
    return CompositeLayer(z3(), z3(), z3(), z3(), z1(), z1())


def rt_kernel_(ctx: _lib.Context, pol_n: int, added_layer: AddedLayer | None, composite_layer: CompositeLayer,
               tau: torch.Tensor, varpi: torch.Tensor, Zpp: torch.Tensor, Zmp: torch.Tensor,
               scattering_interface: int, tau_sum: torch.Tensor, m: int, quad_points: QuadPoints,
               F0: torch.Tensor, iz: int, ndoubl: int, scatter: bool, mu: torch.Tensor, wt: torch.Tensor):
    """One layer of the adding-doubling loop, in place on ``composite_layer``.

    ``ndoubl`` / ``scatter`` are the full-band values of ``get_dtau_ndoubl`` /
    ``maximum(τ.*ϖ) > 2eps`` (rt_kernel.jl:197,266-287); ``Zpp``/``Zmp`` are
    ``(N, N)`` (spectrally constant) or ``(nSpec, N, N)``; ``iz`` is 1-based."""
    nSpec = tau.shape[0]
    N = mu.shape[0]
    d = _lib.LayerDesc()
    d.struct_size = C.sizeof(_lib.LayerDesc)
    d.N, d.nStokes, d.nSpec, d.m, d.imu0 = N, pol_n, nSpec, m, quad_points.imu0
    d.ndoubl, d.scatter, d.iface, d.iz = int(ndoubl), int(bool(scatter)), int(scattering_interface), int(iz)
    d.z_is_spectral = int(Zpp.dim() == 3 and Zpp.shape[0] > 1)
    d.mu0 = float(quad_points.mu0)
    for name, t in (("tau", tau), ("varpi", varpi), ("Zpp", Zpp), ("Zmp", Zmp), ("tau_sum", tau_sum), ("F0", F0),
                    ("mu", mu), ("wt", wt)):
        if not (t.is_cuda and t.is_contiguous()):
            raise ValueError(f"{name}: contiguous CUDA tensor required")
        setattr(d, name, t.data_ptr())
    cl = composite_layer
    d.Rmp, d.Rpm, d.Tpp, d.Tmm, d.J0p, d.J0m = (x.data_ptr() for x in (cl.Rmp, cl.Rpm, cl.Tpp, cl.Tmm, cl.J0p, cl.J0m))
    if added_layer is not None:
        al = added_layer
        d.rmp, d.rpm, d.tpp, d.tmm, d.j0p, d.j0m = (x.data_ptr() for x in (al.rmp, al.rpm, al.tpp, al.tmm, al.j0p, al.j0m))
    fn = _lib.lib.vsm_rt_layer_f64 if tau.dtype == torch.float64 else _lib.lib.vsm_rt_layer_f32
    ctx.check(fn(ctx.handle, C.byref(d)))

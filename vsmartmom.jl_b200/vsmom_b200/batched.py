"""``batched_mul`` / ``batch_inv!`` / ``batch_solve!`` with the semantics of the
reference's CUDA methods (ext/gpu_batched_cuda.jl:72-233) on torch CUDA tensors.

Tensors use the memory order of Julia's column-major (N, N, nSpec) arrays, i.e.
torch shape (nSpec, cols, rows) contiguous: ``A[k, j, i]`` is Julia ``A[i, j, k]``.
"""
from __future__ import annotations

import torch

from . import _lib


def _sfx(t: torch.Tensor) -> str:
    if t.dtype == torch.float64:
        return "f64"
    if t.dtype == torch.float32:
        return "f32"
    raise TypeError("only float32 / float64 are supported (as the reference's CUDA methods)")


def _chk(*ts):
    for t in ts:
        if not (t.is_cuda and t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors (there is no CPU fallback)")


def batched_mul(ctx: _lib.Context, A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """C[:,:,k] = A[:,:,k] * B[:,:,k]; returns a fresh tensor like the reference
    (gpu_batched_cuda.jl:208-218).  A: (batch, K, M), B: (batch, Nc, K) in torch shape."""
    _chk(A, B)
    batch, K, M = A.shape
    b2, Nc, K2 = B.shape
    if b2 != batch or K2 != K:
        raise ValueError(f"batch/inner dim mismatch: {tuple(A.shape)} vs {tuple(B.shape)}")
    Cm = torch.empty((batch, Nc, M), dtype=A.dtype, device=A.device)
    fn = getattr(_lib.lib, f"vsm_batched_mul_{_sfx(A)}")
    ctx.check(fn(ctx.handle, Cm.data_ptr(), A.data_ptr(), B.data_ptr(), M, K, Nc, batch))
    return Cm


def batch_inv_(ctx: _lib.Context, X: torch.Tensor, A: torch.Tensor, info: torch.Tensor | None = None) -> torch.Tensor:
    """X[:,:,k] = inv(A[:,:,k]) in place into X; A is clobbered like getrf
    (gpu_batched_cuda.jl:97-182).  ``info`` (int32, batch) receives 0 or the 1-based zero-pivot index."""
    _chk(X, A)
    batch, N, N2 = A.shape
    if N != N2 or X.shape != A.shape:
        raise ValueError("batch_inv!: square matrices of equal shape required")
    fn = getattr(_lib.lib, f"vsm_batch_inv_{_sfx(A)}")
    ctx.check(fn(ctx.handle, X.data_ptr(), A.data_ptr(), N, batch, info.data_ptr() if info is not None else None))
    return X


def batch_solve_(ctx: _lib.Context, X: torch.Tensor, A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """X[:,:,k] = A[:,:,k] \\ B[:,:,k] (gpu_batched_cuda.jl:72-94: inverse then product)."""
    _chk(X, A, B)
    batch, N, _ = A.shape
    Nc = B.shape[1]
    fn = getattr(_lib.lib, f"vsm_batch_solve_{_sfx(A)}")
    ctx.check(fn(ctx.handle, X.data_ptr(), A.data_ptr(), B.data_ptr(), N, Nc, batch))
    return X

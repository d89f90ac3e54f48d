"""Spectral-axis sharding across ranks (one process per GPU).

The elastic forward path has no cross-spectral term, so the band
is cut into contiguous shards, every rank runs its shard with no data-path
collective, and ONE gather of the (nVZA, pol_n, nSpec_shard) outputs closes the
run.  ndoubl / scatter / ScatteringInterface come from reductions over the FULL
band (``model.layer_control`` before sharding) so every shard branches alike."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(nSpec: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) per rank; interior boundaries are multiples of ``align``."""
    per = -(-nSpec // world)
    per = -(-per // align) * align
    out = []
    for r in range(world):
        lo = min(nSpec, r * per)
        out.append((lo, min(nSpec, lo + per)))
    return out


def shard_bounds_weighted(weights, world: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) per rank with (nearly) equal total WEIGHT instead of equal length -- the Raman path, where the
    work of an output wavelength is its number of in-band donor wavelengths (fewer near the band edges)."""
    import numpy as np
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    c = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(c, c[-1] * r / world, side="left")))
    cuts.append(n)
    cuts = [min(max(x, 0), n) for x in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_spectral(local: torch.Tensor, nSpec: int, world: int, rank: int, align: int = 32, bounds=None) -> torch.Tensor:
    """All-gather shards whose FIRST dimension is the spectral axis into the full
    band (every rank gets the result).  Shards may be ragged / empty.  ``bounds``: explicit [lo, hi) per rank
    (default: ``shard_bounds``)."""
    if world == 1:
        return local
    if bounds is None:
        bounds = shard_bounds(nSpec, world, align)
    per = max(hi - lo for lo, hi in bounds)
    pad = local.new_zeros((per,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    full = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(full, pad)
    parts = [full[r * per: r * per + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return torch.cat(parts, dim=0)

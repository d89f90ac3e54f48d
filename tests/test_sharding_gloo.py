"""world_size-2 gloo test (CPU) of the multi-rank host logic: shard bounds, the
ragged all-gather, and that shards run with FULL-band control flow reproduce the
full band (the oracle stands in for the per-rank GPU call)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vsmom_b200 as vb
from vsmom_b200.sharding import gather_spectral, shard_bounds


def test_shard_bounds_cover_the_band():
    for n, w in [(10000, 8), (22669, 8), (1, 2), (0, 2), (33, 2), (64, 4), (100, 3)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and len(b) == w
        for (l0, h0), (l1, h1) in zip(b, b[1:]):
            assert h0 == l1 and l0 <= h0
        for lo, hi in b[:-1]:
            assert hi == n or hi % 32 == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nspec, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ("vsmartmom.jl_b200", "oracle", "tests"):
        sys.path.insert(0, os.path.join(root, p))
    import vsmom_oracle as orc
    from helpers import to_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = vb.synthetic_o2a_model(nSpec=nspec, nstreams=3, aerosol=True, polarization_type="Stokes_IQU")
    om = to_oracle(m)
    lo, hi = shard_bounds(nspec, world)[rank]
    if hi > lo:
        o = orc.rt_run(om, slice(lo, hi))                  # full-band control flow inside
        local = torch.from_numpy(np.ascontiguousarray(o["R_SFI"].transpose(2, 1, 0)))
    else:
        local = torch.zeros((0, m.pol_n, len(m.vza)), dtype=torch.float64)
    full = gather_spectral(local, nspec, world, rank)
    if rank == 0:
        ref = orc.rt_run(om)["R_SFI"].transpose(2, 1, 0)
        q.put(float(np.max(np.abs(full.numpy() - ref))) if full.shape[0] == nspec else 1e9)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nspec", [70, 20])
def test_two_rank_gather_equals_full_band(nspec):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nspec, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err == 0.0


def _worker_rrs(rank, world, port, nspec, q):
    """Raman owner-range sharding: a rank owns the outputs of a contiguous n1 range and needs only the
    ELASTIC state of its halo; the oracle's point-wise evaluation stands in for the per-rank GPU call."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ("vsmartmom.jl_b200", "oracle", "tests"):
        sys.path.insert(0, os.path.join(root, p))
    import vsmom_oracle as orc
    import vsmom_oracle_rrs as orr
    from helpers import to_oracle
    from vsmom_b200 import raman
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = raman.synthetic_rrs_model(nSpec=nspec, Nz=2)
    rs = raman.synthetic_rrs(grid_step=4.0)
    g = rs.greek_raman
    ors = orr.RRS(np.asarray(rs.i_shift), np.asarray(rs.varpi_raman),
                  orc.GreekCoefs(*(np.asarray(x, float) for x in (g.alpha, g.beta, g.gamma, g.delta, g.epsilon, g.zeta))))
    om = to_oracle(m)
    lo, hi = shard_bounds(nspec, world)[rank]
    k = m.pol_n * len(m.vza)
    if hi > lo:
        o = orr.rt_run_rrs_points(om, ors, np.arange(lo, hi))
        local = torch.from_numpy(np.concatenate([o[key].transpose(2, 1, 0).reshape(hi - lo, k)
                                                 for key in ("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI")], axis=1))
    else:
        local = torch.zeros((0, 4 * k), dtype=torch.float64)
    full = gather_spectral(local, nspec, world, rank).numpy()
    if rank == 0:
        ref = orr.rt_run_rrs(om, ors)
        err = 0.0
        for i, key in enumerate(("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI")):
            got = full[:, i * k:(i + 1) * k].reshape(-1, m.pol_n, len(m.vza)).transpose(2, 1, 0)
            err = max(err, float(np.max(np.abs(got - ref[key])) / np.max(np.abs(ref[key]))))
        q.put(err if full.shape[0] == nspec else 1e9)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_raman_owner_shards_equal_full_band():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rrs, args=(r, 2, port, 40, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-11

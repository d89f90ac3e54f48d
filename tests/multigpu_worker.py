"""torchrun worker for tests/test_multigpu.py: rt_run_sharded over NCCL must equal the
single-GPU full-band run bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vsmartmom.jl_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch
import torch.distributed as dist

import vsmom_b200 as vb
from vsmom_b200 import _lib, rt_run as rr

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nspec = int(sys.argv[1])
mode = sys.argv[2] if len(sys.argv) > 2 else "elastic"
ctx = _lib.Context(local)
ok = True
if mode == "rrs":
    # rotational Raman: owner-range shards with a locally recomputed elastic halo, one all-gather
    from vsmom_b200 import raman
    m = raman.synthetic_rrs_model(nSpec=nspec)
    rs = raman.synthetic_rrs(grid_step=2.0)
    out = raman.rt_run_rrs_sharded(m, rs, ctx, device=f"cuda:{local}")
    if dist.get_rank() == 0:
        ref = raman.rt_run_rrs(m, rs, ctx, device=f"cuda:{local}")
        for k in ("R_SFI", "T_SFI", "ieR_SFI", "ieT_SFI"):
            ok = ok and np.array_equal(out[k], ref[k])
        print("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH", flush=True)
else:
    m = vb.synthetic_o2a_model(nSpec=nspec, nstreams=4, aerosol=True)
    out = rr.rt_run_sharded(m, ctx, device=f"cuda:{local}")
    if dist.get_rank() == 0:
        ref = rr.rt_run(m, ctx, device=f"cuda:{local}")
        for k in ("R_SFI", "T_SFI", "hdr", "bhr_uw", "bhr_dw"):
            ok = ok and np.array_equal(out[k], ref[k])
        print("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH", flush=True)
ctx.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

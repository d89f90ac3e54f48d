"""The C-ABI library loads and exports every symbol include/vsmom_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vsmom_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vsm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("vsm_create", "vsm_batched_mul_f64", "vsm_batch_inv_f32", "vsm_rt_layer_f64", "vsm_rt_column_f64"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from vsmom_b200 import _lib
    names = declared_functions()
    assert names, "no declarations parsed"
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert _lib.lib.vsm_version() >= 100


def test_struct_sizes_match_header_layout():
    from vsmom_b200 import _lib
    # 12 int32 + 1 double + 20 pointers ; 10 int32 + 3 double + 20 pointers (16 before surf_rho, 17 before ss_correction, 18 before thermal_B)
    assert ctypes.sizeof(_lib.LayerDesc) == 12 * 4 + 8 + 20 * 8
    assert ctypes.sizeof(_lib.ColumnDesc) == 10 * 4 + 3 * 8 + 20 * 8
    assert _lib.ColumnDesc.thermal_B.offset == 10 * 4 + 3 * 8 + 18 * 8     # the pre-sources descriptor size, still accepted
    assert _lib.ColumnDesc.surf_rho.offset == 10 * 4 + 3 * 8 + 16 * 8      # the pre-BRDF descriptor size, still accepted
    assert _lib.ColumnDesc.ss_correction.offset == 10 * 4 + 3 * 8 + 17 * 8  # the pre-Cox-Munk descriptor size, still accepted
    assert ctypes.sizeof(_lib.LinDesc) == 2 * 4 + 4 * 8


def test_no_cpu_fallback_without_gpu():
    import torch
    from vsmom_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.VsmError):
        _lib.Context(0)
    import numpy as np
    import vsmom_b200 as vb
    from vsmom_b200 import rt_run as rr
    with pytest.raises(RuntimeError):
        rr.rt_run(vb.synthetic_o2a_model(nSpec=2, aerosol=False))

"""ctypes binding of libvsmom_b200.so (include/vsmom_b200.h).

There is no CPU fallback: if the shared library is missing this module raises
at import, and ``Context()`` raises if no CUDA device is present."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSMOM_B200_LIB", os.path.join(_HERE, "..", "lib", "libvsmom_b200.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"libvsmom_b200.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  The B200 path has no CPU fallback.")
lib = C.CDLL(os.path.abspath(LIB_PATH))

c_void_p, c_int, c_double, c_float = C.c_void_p, C.c_int, C.c_double, C.c_float
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class LayerDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("N", C.c_int32), ("nStokes", C.c_int32), ("nSpec", C.c_int32),
                ("m", C.c_int32), ("imu0", C.c_int32), ("ndoubl", C.c_int32), ("scatter", C.c_int32),
                ("iface", C.c_int32), ("iz", C.c_int32), ("z_is_spectral", C.c_int32), ("reserved", C.c_int32),
                ("mu0", c_double),
                ("tau", c_void_p), ("varpi", c_void_p), ("Zpp", c_void_p), ("Zmp", c_void_p),
                ("tau_sum", c_void_p), ("F0", c_void_p), ("mu", c_void_p), ("wt", c_void_p),
                ("Rmp", c_void_p), ("Rpm", c_void_p), ("Tpp", c_void_p), ("Tmm", c_void_p),
                ("J0p", c_void_p), ("J0m", c_void_p),
                ("rmp", c_void_p), ("rpm", c_void_p), ("tpp", c_void_p), ("tmm", c_void_p),
                ("j0p", c_void_p), ("j0m", c_void_p)]


class ColumnDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("N", C.c_int32), ("nStokes", C.c_int32), ("nSpec", C.c_int32),
                ("Nz", C.c_int32), ("nAer", C.c_int32), ("nVZA", C.c_int32), ("n_moments", C.c_int32),
                ("imu0", C.c_int32), ("surface_kind", C.c_int32),
                ("mu0", c_double), ("albedo", c_double), ("varpi_cabannes", c_double),
                ("mu", f64p), ("wt", f64p), ("aer_omega", f64p), ("aer_ft", f64p), ("tau_aer", f64p),
                ("vza_weight", f64p),
                ("ndoubl", i32p), ("scatter", i32p), ("iface", i32p), ("rayl_allzero", i32p), ("vza_stream", i32p),
                ("tau_rayl", c_void_p), ("tau_abs", c_void_p), ("F0", c_void_p), ("Zpp", c_void_p), ("Zmp", c_void_p),
                ("surf_rho", c_void_p), ("ss_correction", f64p),
                ("thermal_B", c_void_p), ("sif0", c_void_p)]


class LinDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_params", C.c_int32), ("dtau_aer", f64p), ("dalbedo", f64p),
                ("dtau_rayl", c_void_p), ("dtau_abs", c_void_p)]


class RrsDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_raman", C.c_int32), ("n1_begin", C.c_int32), ("n1_count", C.c_int32),
                ("i_shift", i32p), ("varpi_raman", f64p), ("Zpp_raman", c_void_p), ("Zmp_raman", c_void_p)]


EXPORTS = [
    "vsm_version", "vsm_create", "vsm_destroy", "vsm_last_error", "vsm_synchronize",
    "vsm_set_workspace_limit", "vsm_set_inverse_mode", "vsm_kernel_launches", "vsm_measure_fma_peak",
    "vsm_batched_mul_f64", "vsm_batched_mul_f32", "vsm_batch_inv_f64", "vsm_batch_inv_f32",
    "vsm_batch_solve_f64", "vsm_batch_solve_f32",
    "vsm_rt_layer_f64", "vsm_rt_layer_f32", "vsm_rt_column_f64", "vsm_rt_column_f32",
    "vsm_rt_column_streams_f64", "vsm_rt_column_streams_f32", "vsm_rt_column_lin_f64", "vsm_rt_column_rrs_f64", "vsm_rt_column_ss_f64", "vsm_rt_column_ss_f32", "vsm_rt_column_ss_rrs_f64",
    "vsm_last_column_ms", "vsm_last_column_launches", "vsm_last_column_status",
    "vsm_multi_create", "vsm_multi_destroy", "vsm_multi_last_error", "vsm_multi_ndev", "vsm_multi_ctx",
    "vsm_rt_column_sharded_f64", "vsm_rt_column_sharded_f32", "vsm_rt_column_rrs_sharded_f64",
]

lib.vsm_version.restype = c_int
lib.vsm_create.argtypes = [C.POINTER(c_void_p), c_int, c_void_p]
lib.vsm_destroy.argtypes = [c_void_p]
lib.vsm_last_error.argtypes = [c_void_p]
lib.vsm_last_error.restype = C.c_char_p
lib.vsm_synchronize.argtypes = [c_void_p]
lib.vsm_set_workspace_limit.argtypes = [c_void_p, C.c_uint64]
lib.vsm_set_inverse_mode.argtypes = [c_void_p, C.c_int]
lib.vsm_kernel_launches.argtypes = [c_void_p]
lib.vsm_kernel_launches.restype = C.c_uint64
lib.vsm_measure_fma_peak.argtypes = [c_void_p, c_int, C.POINTER(c_double)]
for _s in ("f64", "f32"):
    getattr(lib, f"vsm_batched_mul_{_s}").argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int]
    getattr(lib, f"vsm_batch_inv_{_s}").argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    getattr(lib, f"vsm_batch_solve_{_s}").argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]
    getattr(lib, f"vsm_rt_layer_{_s}").argtypes = [c_void_p, C.POINTER(LayerDesc)]
    getattr(lib, f"vsm_rt_column_{_s}").argtypes = [c_void_p, C.POINTER(ColumnDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    getattr(lib, f"vsm_rt_column_streams_{_s}").argtypes = [c_void_p, C.POINTER(ColumnDesc)] + [c_void_p] * 9
lib.vsm_rt_column_lin_f64.argtypes = [c_void_p, C.POINTER(ColumnDesc), C.POINTER(LinDesc)] + [c_void_p] * 4
lib.vsm_rt_column_rrs_f64.argtypes = [c_void_p, C.POINTER(ColumnDesc), C.POINTER(RrsDesc)] + [c_void_p] * 4
for _s in ("f64", "f32"):
    getattr(lib, f"vsm_rt_column_ss_{_s}").argtypes = [c_void_p, C.POINTER(ColumnDesc)] + [c_void_p] * 4
lib.vsm_last_column_ms.argtypes = [c_void_p, C.POINTER(c_float), C.POINTER(c_float), C.POINTER(c_float)]
lib.vsm_last_column_launches.argtypes = [c_void_p, C.POINTER(c_int), C.POINTER(c_int)]
lib.vsm_last_column_status.argtypes = [c_void_p, C.POINTER(c_int)]


lib.vsm_multi_create.argtypes = [C.POINTER(c_void_p), i32p, c_int]
lib.vsm_multi_destroy.argtypes = [c_void_p]
lib.vsm_multi_last_error.argtypes = [c_void_p]
lib.vsm_multi_last_error.restype = C.c_char_p
lib.vsm_multi_ndev.argtypes = [c_void_p]
lib.vsm_multi_ctx.argtypes = [c_void_p, c_int]
lib.vsm_multi_ctx.restype = c_void_p
for _s in ("f64", "f32"):
    getattr(lib, f"vsm_rt_column_sharded_{_s}").argtypes = [c_void_p, C.POINTER(ColumnDesc)] + [c_void_p] * 5

lib.vsm_rt_column_ss_rrs_f64.argtypes = [c_void_p, C.POINTER(ColumnDesc), C.POINTER(RrsDesc)] + [c_void_p] * 6
lib.vsm_rt_column_rrs_sharded_f64.argtypes = [c_void_p, C.POINTER(ColumnDesc), C.POINTER(RrsDesc)] + [c_void_p] * 4


class VsmError(RuntimeError):
    pass


class Context:
    """One ``vsm_ctx`` (device + stream).  Mirrors what ``vSmartMOMCUDAExt.__init__``
    sets up (ext/vSmartMOMCUDAExt.jl:36-66): a functional-device check and the
    synchronisation hook."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = c_void_p()
        rc = lib.vsm_create(C.byref(h), int(device), c_void_p(stream) if stream else None)
        if rc != 0:
            raise VsmError(f"vsm_create failed ({rc}): {lib.vsm_last_error(None).decode()}")
        self.handle = h
        self.device = device

    def check(self, rc: int):
        if rc != 0:
            raise VsmError(f"libvsmom_b200 error {rc}: {lib.vsm_last_error(self.handle).decode()}")

    def synchronize(self):
        self.check(lib.vsm_synchronize(self.handle))

    def set_workspace_limit(self, nbytes: int):
        self.check(lib.vsm_set_workspace_limit(self.handle, int(nbytes)))

    def set_inverse_mode(self, mode: int):
        """1: adaptive Neumann / Gauss-Jordan (default); 0: Gauss-Jordan only."""
        self.check(lib.vsm_set_inverse_mode(self.handle, int(mode)))

    @property
    def kernel_launches(self) -> int:
        return int(lib.vsm_kernel_launches(self.handle))

    def measure_fma_peak(self, kind: int) -> float:
        """TFLOP/s of register-resident FMA chains: kind 0 = fp64, 1 = fp32."""
        v = c_double()
        self.check(lib.vsm_measure_fma_peak(self.handle, int(kind), C.byref(v)))
        return v.value

    def last_column_ms(self):
        a, b, c = c_float(), c_float(), c_float()
        self.check(lib.vsm_last_column_ms(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def last_column_launches(self):
        a, b = c_int(), c_int()
        self.check(lib.vsm_last_column_launches(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_column_nonfinite(self) -> int:
        """Number of non-finite R_SFI / T_SFI elements of the most recent column call (0 = healthy)."""
        v = c_int()
        self.check(lib.vsm_last_column_status(self.handle, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "handle", None):
            lib.vsm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext:
    """``vsm_multi``: several GPUs driven from this one process (in-library spectral sharding + one NCCL all-gather)."""

    def __init__(self, devices):
        import numpy as np
        dv = np.ascontiguousarray(list(devices), np.int32)
        h = c_void_p()
        rc = lib.vsm_multi_create(C.byref(h), dv.ctypes.data_as(i32p), int(dv.size))
        if rc != 0:
            raise VsmError(f"vsm_multi_create failed ({rc}): {lib.vsm_multi_last_error(None).decode()}")
        self.handle = h
        self.devices = list(map(int, dv))

    def check(self, rc: int):
        if rc != 0:
            raise VsmError(f"libvsmom_b200 error {rc}: {lib.vsm_multi_last_error(self.handle).decode()}")

    def close(self):
        if getattr(self, "handle", None):
            lib.vsm_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

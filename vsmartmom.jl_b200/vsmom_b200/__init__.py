"""vsmom_b200 -- host-side mirror of the reference interface for the CoreRT hot
path, on top of the sm_100a CUDA library (C-ABI: include/vsmom_b200.h)."""
from .model import (AerosolOptics, GreekCoefs, QuadPoints, RTModel, RTNumericalParameters, compute_Z_moments,
                    derive_m_max, get_greek_rayleigh, hg_greek, layer_control, brdf_fourier_moments, rossli_surface_rho, rpv_surface_rho, rt_set_streams,
                    synthetic_o2a_model, synthetic_ocean_model)
from .coxmunk import CoxMunkSurface, coxmunk_ss_correction, coxmunk_surface_rho

__all__ = ["AerosolOptics", "GreekCoefs", "QuadPoints", "RTModel", "RTNumericalParameters", "compute_Z_moments",
           "derive_m_max", "get_greek_rayleigh", "hg_greek", "layer_control", "brdf_fourier_moments", "rossli_surface_rho", "rpv_surface_rho", "rt_set_streams",
           "synthetic_o2a_model", "synthetic_ocean_model", "CoxMunkSurface", "coxmunk_ss_correction", "coxmunk_surface_rho"]

"""ctypes binding of libbarcode_b200.so, the C ABI declared in include/barcode_gpu.h.

There is no CPU fallback anywhere in this package: if the shared library is
missing, or no CUDA device is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbarcode_b200.so")

BGPU_CALC_H_EXACT = 4


class BgpuParams(C.Structure):
    """`bgpu_params` (include/barcode_gpu.h)."""
    _fields_ = [
        ("N1", C.c_int), ("N2", C.c_int), ("N3", C.c_int),
        ("L1", C.c_double), ("L2", C.c_double), ("L3", C.c_double),
        ("xllc", C.c_double), ("yllc", C.c_double), ("zllc", C.c_double),
        ("xobs", C.c_double), ("yobs", C.c_double), ("zobs", C.c_double),
        ("planepar", C.c_int), ("periodic", C.c_int),
        ("masskernel", C.c_int), ("likelihood", C.c_int), ("sfmodel", C.c_int), ("rsd_model", C.c_int),
        ("calc_h", C.c_int), ("mass_type", C.c_int),
        ("D1", C.c_double), ("D2", C.c_double), ("ascale", C.c_double), ("OM", C.c_double), ("OL", C.c_double),
        ("particle_kernel_h_rel", C.c_double), ("slength", C.c_double),
        ("rho_c", C.c_double), ("biasP", C.c_double), ("biasE", C.c_double),
        ("deltaQ_factor", C.c_double), ("correct_delta", C.c_int),
        ("mass_factor", C.c_double),
        ("div_dH_by_N", C.c_int),
        ("device", C.c_int),
        ("delta_min", C.c_double),
        ("N_bin", C.c_int),
        ("reserved", C.c_int * 5),
    ]


# every symbol include/barcode_gpu.h declares: (restype, argtypes)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
_h = C.c_void_p
SIGNATURES = {
    "bgpu_default_params": (None, [C.POINTER(BgpuParams)]),
    "bgpu_abi_version": (C.c_int, []),
    "bgpu_last_error": (C.c_char_p, []),
    "bgpu_create": (C.c_int, [C.POINTER(BgpuParams), C.POINTER(_h)]),
    "bgpu_destroy": (None, [_h]),
    "bgpu_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "bgpu_slab_create": (C.c_int, [C.POINTER(BgpuParams), C.c_int, C.c_int, C.c_void_p, C.POINTER(_h)]),
    "bgpu_slab_info": (C.c_int, [_h, _ip, _ip, _ip, _ip]),
    "bgpu_set_static": (C.c_int, [_h, _dp, _dp, _dp, _dp]),
    "bgpu_set_mass": (C.c_int, [_h, _dp, _dp]),
    "bgpu_hamiltonian_mass": (C.c_int, [_h, _dp, _dp]),
    "bgpu_gradient_psi": (C.c_int, [_h, _dp, _dp]),
    "bgpu_psi": (C.c_int, [_h, _dp, _dp, _dp, _dp]),
    "bgpu_kinetic": (C.c_int, [_h, _dp, _dp]),
    "bgpu_leapfrog": (C.c_int, [_h, _dp, _dp, C.c_uint64, C.c_double, _dp, _dp]),
    "bgpu_color_momenta": (C.c_int, [_h, _dp, _dp, _dp]),
    "bgpu_draw_momenta_device": (C.c_int, [_h, C.c_uint64, C.c_uint64, _dp]),
    "bgpu_draw_momenta_device_dev": (C.c_int, [_h, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bgpu_hamiltonian_mass_x": (C.c_int, [_h, _dp, _dp, _dp]),
    "bgpu_likeli_force_power": (C.c_int, [_h, _dp, _dp, _dp]),
    "bgpu_measure_spectrum": (C.c_int, [_h, _dp, C.c_uint64, _dp, _dp]),
    "bgpu_set_signal": (C.c_int, [_h, _dp]),
    "bgpu_candidate": (C.c_int, [_h, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, _dp, _dp]),
    "bgpu_accept": (C.c_int, [_h, _dp, _dp]),
    "bgpu_device_normals": (C.c_int, [_h, C.c_uint64, C.c_uint64, C.c_uint, C.c_size_t, C.c_size_t, _dp]),
    "bgpu_local_group_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "bgpu_local_group_destroy": (None, [C.c_void_p]),
    "bgpu_slab_create_local": (C.c_int, [C.POINTER(BgpuParams), C.c_int, C.c_int, C.c_void_p, C.POINTER(_h)]),
    "bgpu_mock_data": (C.c_int, [_h, C.c_uint64, C.c_void_p, _dp, _dp, _dp, _dp, _dp]),
    "bgpu_initial_guess": (C.c_int, [_h, C.c_uint64, C.c_int, C.c_double, _dp]),
    "bgpu_forward": (C.c_int, [_h, _dp, _dp, _dp, _dp, _dp]),
    "bgpu_assign_density": (C.c_int, [_h, _dp, _dp, _dp, _dp]),
    "bgpu_cell_indices": (C.c_int, [_h, _dp, _dp, _dp, C.c_size_t, _ip, _ip, _ip]),
    "bgpu_fft_r2c": (C.c_int, [_h, _dp, _dp]),
    "bgpu_fft_c2r": (C.c_int, [_h, _dp, _dp]),
    "bgpu_convolve_inv_corr": (C.c_int, [_h, _dp, _dp, _dp]),
    "bgpu_set_stream": (C.c_int, [_h, C.c_void_p]),
    "bgpu_synchronize": (C.c_int, [_h]),
    "bgpu_gradient_psi_dev": (C.c_int, [_h, C.c_void_p, C.c_void_p]),
    "bgpu_psi_dev": (C.c_int, [_h, C.c_void_p, _dp, _dp, C.c_void_p]),
    "bgpu_kinetic_dev": (C.c_int, [_h, C.c_void_p, _dp]),
    "bgpu_leapfrog_dev": (C.c_int, [_h, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double]),
    # single-precision mode (reference build option SINGLE_PREC)
    "bgpu_f32_create": (C.c_int, [C.POINTER(BgpuParams), C.POINTER(_h)]),
    "bgpu_f32_destroy": (None, [_h]),
    "bgpu_f32_set_static": (C.c_int, [_h, _fp, _fp, _fp, _fp]),
    "bgpu_f32_set_mass": (C.c_int, [_h, _fp, _fp]),
    "bgpu_f32_hamiltonian_mass": (C.c_int, [_h, _fp, _fp]),
    "bgpu_f32_gradient_psi": (C.c_int, [_h, _fp, _fp]),
    "bgpu_f32_psi": (C.c_int, [_h, _fp, _dp, _dp, _fp]),
    "bgpu_f32_kinetic": (C.c_int, [_h, _fp, _dp]),
    "bgpu_f32_leapfrog": (C.c_int, [_h, _fp, _fp, C.c_uint64, C.c_double, _fp, _fp]),
    "bgpu_f32_set_stream": (C.c_int, [_h, C.c_void_p]),
    "bgpu_f32_synchronize": (C.c_int, [_h]),
    "bgpu_f32_gradient_psi_dev": (C.c_int, [_h, C.c_void_p, C.c_void_p]),
    "bgpu_f32_psi_dev": (C.c_int, [_h, C.c_void_p, _dp, _dp, C.c_void_p]),
    "bgpu_f32_leapfrog_dev": (C.c_int, [_h, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double]),
    "bgpu_kernel_launches": (C.c_uint64, []),
    "bgpu_profile_begin": (C.c_int, []),
    "bgpu_profile_end": (C.c_int, [_dp, C.POINTER(C.c_uint64), C.c_int]),
    "bgpu_profile_kind_name": (C.c_char_p, [C.c_int]),
}
PROFILE_KINDS = 15

_lib = None


def load():
    """Load the shared library (once) and type every entry point; raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m barcode_b200.build` "
                "(there is no CPU fallback for this path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class BgpuError(RuntimeError):
    """What the C++ glue rethrows as std::runtime_error (main.cc:195-197)."""


def check(rc: int):
    if rc != 0:
        raise BgpuError(load().bgpu_last_error().decode(errors="replace"))

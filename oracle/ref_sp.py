"""oracle/ref_sp.py -- TEST INFRASTRUCTURE, not product code.

ctypes binding of ``oracle/_ref/libbarcode_ref_sp.so``: the unmodified Barcode reference sources compiled with the
reference's own SINGLE_PREC option (real_prec = float; define_opt.h:50-59, cmake/Modules/Options.cmake:65-66) against
the shim headers, with ``fftwf_*`` provided by ``oracle/shim/fftw_shim.cc`` (widen, transform in double, round once).
It is the checker of the single-precision mode (``bgpu_f32_*``): only ``tests/`` may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .ref import Config, RefParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbarcode_ref_sp.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} not built: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.POINTER(RefParams)]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_last_error.restype = C.c_char_p
        L.ref_array.restype = fp
        L.ref_array.argtypes = [C.c_void_p, C.c_char_p]
        for name, args in {
            "ref_gradient_psi": [C.c_void_p, fp, fp],
            "ref_psi": [C.c_void_p, fp, dp, dp],
            "ref_kinetic": [C.c_void_p, fp, dp],
            "ref_EoM": [C.c_void_p, fp, fp, fp, fp, C.c_double, C.c_double],
            "ref_hamiltonian_mass": [C.c_void_p],
        }.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32).ravel()


def _chk(rc):
    if rc != 0:
        raise RuntimeError("reference (SINGLE_PREC) threw: " + lib().ref_last_error().decode())


class ReferenceSP:
    """One `DATA` + `HAMIL_DATA` pair of the reference compiled with SINGLE_PREC; arrays are float32."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        p = RefParams()
        for f, _ in RefParams._fields_:
            setattr(p, f, type(getattr(p, f))(getattr(cfg, f)))
        self.h = lib().ref_create(C.byref(p))
        if not self.h:
            raise RuntimeError("ref_create failed: " + lib().ref_last_error().decode())
        self.N1 = cfg.N1
        self.N = cfg.N1 ** 3

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def array(self, name: str) -> np.ndarray:
        ptr = lib().ref_array(self.h, name.encode())
        if not ptr:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, shape=(self.N,))

    def set_inputs(self, Power=None, nobs=None, noise=None, window=None, signal=None):
        for name, val in (("Power", Power), ("nobs", nobs), ("noise", noise), ("window", window), ("signal", signal)):
            if val is not None:
                self.array(name)[:] = _f32(val)

    def hamiltonian_mass(self):
        _chk(lib().ref_hamiltonian_mass(self.h))
        return self.array("mass_f").copy(), self.array("mass_r").copy()

    def gradient_psi(self, s):
        s = _f32(s)
        out = np.empty(self.N, dtype=np.float32)
        _chk(lib().ref_gradient_psi(self.h, _p(s), _p(out)))
        return out

    def psi(self, s):
        s = _f32(s)
        a, b = C.c_double(), C.c_double()
        _chk(lib().ref_psi(self.h, _p(s), C.byref(a), C.byref(b)))
        return a.value, b.value

    def kinetic(self, p):
        p = _f32(p)
        k = C.c_double()
        _chk(lib().ref_kinetic(self.h, _p(p), C.byref(k)))
        return k.value

    def EoM(self, si, pi, u_Neps, u_eps):
        si, pi = _f32(si), _f32(pi)
        sf, pf = np.empty(self.N, dtype=np.float32), np.empty(self.N, dtype=np.float32)
        _chk(lib().ref_EoM(self.h, _p(si), _p(pi), _p(sf), _p(pf), u_Neps, u_eps))
        return sf, pf

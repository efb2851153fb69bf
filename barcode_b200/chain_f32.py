"""`ChainF32`: the single-precision mode of one HMC chain (`bgpu_f32_*`, include/barcode_gpu.h).

The reference selects its arithmetic at build time (SINGLE_PREC: real_prec = float, define_opt.h:50-59); a
SINGLE_PREC build binds these entry points where a DOUBLE_PREC build binds `Chain`'s.  Arrays are numpy float32,
energies Python floats (double).  Scope: Zel'dovich + CIC (+ plane-parallel RSD), Poisson / Gaussian likelihood,
calc_h 0 / 1 / 4, mass_type 0 / 1 / 4, N1 = 32 ... 512, one GPU -- anything else raises `BgpuError`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .chain import Params


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} elements, got {a.size}")
    return a


class ChainF32:
    def __init__(self, params: Params):
        self.params = params
        self.L = _lib.load()
        self.N1 = int(params.N1)
        self.N = self.N1 ** 3
        self._h = C.c_void_p()
        cp = params.to_c()
        _lib.check(self.L.bgpu_f32_create(C.byref(cp), C.byref(self._h)))

    def close(self):
        if self._h:
            self.L.bgpu_f32_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def shape(self):
        return (self.N1, self.N1, self.N1)

    def set_static(self, Power=None, nobs=None, noise=None, window=None):
        arrs = [None if a is None else _f32(a, self.N) for a in (Power, nobs, noise, window)]
        _lib.check(self.L.bgpu_f32_set_static(self._h, *[None if a is None else _fp(a) for a in arrs]))

    def set_mass(self, mass_f=None, mass_r=None):
        arrs = [None if a is None else _f32(a, self.N) for a in (mass_f, mass_r)]
        _lib.check(self.L.bgpu_f32_set_mass(self._h, *[None if a is None else _fp(a) for a in arrs]))

    def hamiltonian_mass(self):
        """Hamiltonian_mass (HMC_mass.cc:315-368), types 0 / 1 / 4: returns (mass_f, mass_r)."""
        mf, mr = np.zeros(self.N, dtype=np.float32), np.zeros(self.N, dtype=np.float32)
        _lib.check(self.L.bgpu_f32_hamiltonian_mass(self._h, _fp(mf), _fp(mr)))
        return mf.reshape(self.shape), mr.reshape(self.shape)

    def gradient_psi(self, signal):
        """gradient_psi (HMC.cc:146-206)."""
        s = _f32(signal, self.N)
        out = np.empty(self.N, dtype=np.float32)
        _lib.check(self.L.bgpu_f32_gradient_psi(self._h, _fp(s), _fp(out)))
        return out.reshape(self.shape)

    def psi(self, signal):
        """psi (HMC.cc:124-143): (psi_prior, psi_likeli, deltaX)."""
        s = _f32(signal, self.N)
        a, b = C.c_double(), C.c_double()
        dX = np.empty(self.N, dtype=np.float32)
        _lib.check(self.L.bgpu_f32_psi(self._h, _fp(s), C.byref(a), C.byref(b), _fp(dX)))
        return a.value, b.value, dX.reshape(self.shape)

    def kinetic_term(self, momenta):
        """kinetic_term (HMC.cc:64-121)."""
        p = _f32(momenta, self.N)
        k = C.c_double()
        _lib.check(self.L.bgpu_f32_kinetic(self._h, _fp(p), C.byref(k)))
        return k.value

    def leapfrog(self, s_i, p_i, Neps: int, epsilon: float):
        """Hamiltonian_EoM after the RNG draws (HMC.cc:251-369): (s_f, p_f)."""
        s, p = _f32(s_i, self.N), _f32(p_i, self.N)
        sf, pf = np.empty(self.N, dtype=np.float32), np.empty(self.N, dtype=np.float32)
        _lib.check(self.L.bgpu_f32_leapfrog(self._h, _fp(s), _fp(p), int(Neps), float(epsilon), _fp(sf), _fp(pf)))
        return sf.reshape(self.shape), pf.reshape(self.shape)

    # device-pointer variants (bench.py: torch tensors on the chain's stream)
    def set_stream(self, cuda_stream: int):
        _lib.check(self.L.bgpu_f32_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        _lib.check(self.L.bgpu_f32_synchronize(self._h))

    def gradient_psi_dev(self, d_signal: int, d_gradpsi: int):
        _lib.check(self.L.bgpu_f32_gradient_psi_dev(self._h, C.c_void_p(d_signal), C.c_void_p(d_gradpsi)))

    def leapfrog_dev(self, d_signal: int, d_momenta: int, Neps: int, epsilon: float):
        _lib.check(self.L.bgpu_f32_leapfrog_dev(self._h, C.c_void_p(d_signal), C.c_void_p(d_momenta), int(Neps),
                                                 float(epsilon)))

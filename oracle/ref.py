"""oracle/ref.py -- TEST INFRASTRUCTURE, not product code.

ctypes binding of ``oracle/_ref/libbarcode_ref.so``: the unmodified Barcode
reference sources compiled against shim headers (see ``oracle/Makefile`` and
``oracle/ref_harness.cc``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbarcode_ref.so")


class RefParams(C.Structure):
    _fields_ = [
        ("N1", C.c_int),
        ("L1", C.c_double),
        ("xllc", C.c_double), ("yllc", C.c_double), ("zllc", C.c_double),
        ("xobs", C.c_double), ("yobs", C.c_double), ("zobs", C.c_double),
        ("planepar", C.c_int), ("periodic", C.c_int),
        ("masskernel", C.c_int),
        ("likelihood", C.c_int),
        ("sfmodel", C.c_int),
        ("rsd_model", C.c_int),
        ("calc_h", C.c_int),
        ("mass_type", C.c_int),
        ("z", C.c_double),
        ("deltaQ_factor", C.c_double),
        ("correct_delta", C.c_int),
        ("particle_kernel_h_rel", C.c_double),
        ("slength", C.c_double),
        ("N_eps_fac", C.c_double), ("eps_fac", C.c_double),
        ("mass_factor", C.c_double),
        ("div_dH_by_N", C.c_int),
        ("sigma_min", C.c_double), ("sigma_fac", C.c_double), ("delta_min", C.c_double),
        ("N_bin", C.c_int),
    ]


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} not built: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.POINTER(RefParams)]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_last_error.restype = C.c_char_p
        L.ref_array.restype = dp
        L.ref_array.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_scalar.restype = C.c_double
        L.ref_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_fft_backend.restype = C.c_char_p
        for name, args in {
            "ref_readtab": [C.c_void_p, C.c_char_p],
            "ref_gradient_psi": [C.c_void_p, dp, dp],
            "ref_grad_log_like": [C.c_void_p, dp, dp],
            "ref_grad_log_prior": [C.c_void_p, dp, dp],
            "ref_psi": [C.c_void_p, dp, dp, dp],
            "ref_kinetic": [C.c_void_p, dp, dp],
            "ref_EoM": [C.c_void_p, dp, dp, dp, dp, C.c_double, C.c_double],
            "ref_delta_hamiltonian": [C.c_void_p, dp, dp, dp, dp, dp],
            "ref_draw_momenta": [C.c_void_p, C.c_ulong, dp],
            "ref_create_garfield": [C.c_void_p, C.c_ulong, dp, dp],
            "ref_white_noise": [C.c_int, C.c_ulong, dp],
            "ref_rng_stream": [C.c_ulong, C.c_int, C.POINTER(C.c_ulong), C.c_int, dp],
            "ref_hamiltonian_mass": [C.c_void_p],
            "ref_forward": [C.c_void_p, dp, dp, dp, dp, dp],
            "ref_density": [C.c_void_p, dp, dp, dp, dp],
            "ref_partial_f": [C.c_void_p, dp, dp],
            "ref_convolve_inv_corr": [C.c_void_p, dp, dp, dp],
            "ref_fft_r2c": [C.c_int, dp, dp],
            "ref_fft_c2r": [C.c_int, dp, dp],
            "ref_gradfft": [C.c_int, C.c_double, dp, dp, C.c_uint],
            "ref_gradfindif": [C.c_int, C.c_double, dp, dp, C.c_uint],
            "ref_measure_spectrum": [C.c_int, C.c_double, dp, C.c_ulong, dp, dp],
            "ref_time_gradient_psi": [C.c_void_p, dp, C.c_int, dp],
        }.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        L.ref_num_threads.restype = C.c_int
        L.ref_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _chk(rc):
    if rc != 0:
        raise RuntimeError("reference threw: " + lib().ref_last_error().decode())


@dataclass
class Config:
    """Run-time knobs of the reference's input.par that touch the hot path."""
    N1: int = 16
    L1: float = 50.0
    masskernel: int = 1      # 0 NGP, 1 CIC, 2 TSC, 3 SPH
    likelihood: int = 1      # 0 Poisson, 1 Gaussian
    sfmodel: int = 1
    rsd_model: bool = False
    calc_h: int = 0
    mass_type: int = 1
    z: float = 0.0
    deltaQ_factor: float = 1.0
    correct_delta: bool = True
    particle_kernel_h_rel: float = 1.0
    slength: float = 4.0
    N_eps_fac: float = 8.0
    eps_fac: float = 1.0
    mass_factor: float = 1.0
    div_dH_by_N: bool = False
    xllc: float = 0.0
    yllc: float = 0.0
    zllc: float = 0.0
    xobs: float = 90.0
    yobs: float = 90.0
    zobs: float = 90.0
    planepar: bool = True
    periodic: bool = True
    sigma_min: float = 1.0
    sigma_fac: float = 0.0
    delta_min: float = -0.999
    N_bin: int = 200


class Reference:
    """One `DATA` + `HAMIL_DATA` pair of the compiled reference."""

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

    def scalar(self, name: str) -> float:
        return float(lib().ref_scalar(self.h, name.encode()))

    def set_inputs(self, Power=None, nobs=None, noise=None, window=None, signal=None):
        for name, val in (("Power", Power), ("nobs", nobs), ("noise", noise), ("window", window), ("signal", signal)):
            if val is not None:
                self.array(name)[:] = np.asarray(val, dtype=np.float64).ravel()

    def readtab(self, fname: str) -> np.ndarray:
        _chk(lib().ref_readtab(self.h, fname.encode()))
        return self.array("Power").copy()

    def kernelcomp(self):
        """Write <cwd>/auxkernelr<int(slength)>.dat, which the non-Zel'dovich forward model re-reads."""
        _chk(lib().ref_kernelcomp(self.h))

    def hamiltonian_mass(self):
        _chk(lib().ref_hamiltonian_mass(self.h))
        return self.array("mass_f").copy(), self.array("mass_r").copy()

    def gradient_psi(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        out = np.empty(self.N)
        _chk(lib().ref_gradient_psi(self.h, _p(s), _p(out)))
        return out

    def grad_log_like(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        out = np.empty(self.N)
        _chk(lib().ref_grad_log_like(self.h, _p(s), _p(out)))
        return out

    def grad_log_prior(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        out = np.empty(self.N)
        _chk(lib().ref_grad_log_prior(self.h, _p(s), _p(out)))
        return out

    def psi(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        a, b = C.c_double(), C.c_double()
        _chk(lib().ref_psi(self.h, _p(s), C.byref(a), C.byref(b)))
        return a.value, b.value

    def kinetic(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64).ravel()
        k = C.c_double()
        _chk(lib().ref_kinetic(self.h, _p(p), C.byref(k)))
        return k.value

    def EoM(self, si, pi, u_Neps, u_eps):
        si = np.ascontiguousarray(si, dtype=np.float64).ravel()
        pi = np.ascontiguousarray(pi, dtype=np.float64).ravel()
        sf, pf = np.empty(self.N), np.empty(self.N)
        _chk(lib().ref_EoM(self.h, _p(si), _p(pi), _p(sf), _p(pf), u_Neps, u_eps))
        return sf, pf

    def delta_hamiltonian(self, si, pi, sf, pf):
        arrs = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (si, pi, sf, pf)]
        d = C.c_double()
        _chk(lib().ref_delta_hamiltonian(self.h, *[_p(a) for a in arrs], C.byref(d)))
        names = ["dH", "dK", "dE", "dprior", "dlikeli", "psi_prior_i", "psi_prior_f", "psi_likeli_i",
                 "psi_likeli_f", "H_kin_i", "H_kin_f"]
        return d.value, {k: self.scalar(k) for k in names}

    def draw_momenta(self, seed: int):
        out = np.empty(self.N)
        _chk(lib().ref_draw_momenta(self.h, seed, _p(out)))
        return out

    def create_garfield(self, seed: int, power):
        power = np.ascontiguousarray(power, dtype=np.float64).ravel()
        out = np.empty(self.N)
        _chk(lib().ref_create_garfield(self.h, seed, _p(power), _p(out)))
        return out

    def forward(self, s, want_pos=True):
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        d = np.empty(self.N)
        if want_pos:
            x, y, z = np.empty(self.N), np.empty(self.N), np.empty(self.N)
            _chk(lib().ref_forward(self.h, _p(s), _p(d), _p(x), _p(y), _p(z)))
            return d, x, y, z
        _chk(lib().ref_forward(self.h, _p(s), _p(d), None, None, None))
        return d

    def density(self, x, y, z):
        x, y, z = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (x, y, z)]
        rho = np.empty(self.N)
        _chk(lib().ref_density(self.h, _p(x), _p(y), _p(z), _p(rho)))
        return rho

    def partial_f(self, deltaX):
        deltaX = np.ascontiguousarray(deltaX, dtype=np.float64).ravel()
        out = np.empty(self.N)
        _chk(lib().ref_partial_f(self.h, _p(deltaX), _p(out)))
        return out

    def convolve_inv_corr(self, s, corr):
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        corr = np.ascontiguousarray(corr, dtype=np.float64).ravel()
        out = np.empty(self.N)
        _chk(lib().ref_convolve_inv_corr(self.h, _p(s), _p(corr), _p(out)))
        return out

    def time_gradient_psi(self, s, reps: int) -> float:
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        t = C.c_double()
        _chk(lib().ref_time_gradient_psi(self.h, _p(s), reps, C.byref(t)))
        return t.value


def white_noise(N1: int, seed: int) -> np.ndarray:
    out = np.empty(2 * N1 ** 3)
    _chk(lib().ref_white_noise(N1, seed, _p(out)))
    return out.view(np.complex128).reshape(N1, N1, N1)


def rng_stream(seed: int, n_raw: int, n_gauss: int):
    raw = (C.c_ulong * n_raw)()
    g = np.empty(n_gauss)
    lib().ref_rng_stream(seed, n_raw, raw, n_gauss, _p(g))
    return np.array(raw[:], dtype=np.uint64), g


def fft_r2c(a: np.ndarray) -> np.ndarray:
    N1 = a.shape[0]
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty(2 * N1 * N1 * (N1 // 2 + 1))
    _chk(lib().ref_fft_r2c(N1, _p(a.ravel()), _p(out)))
    return out.view(np.complex128).reshape(N1, N1, N1 // 2 + 1)


def fft_c2r(c: np.ndarray) -> np.ndarray:
    N1 = c.shape[0]
    c = np.ascontiguousarray(c, dtype=np.complex128)
    out = np.empty(N1 ** 3)
    _chk(lib().ref_fft_c2r(N1, _p(c.view(np.float64).ravel()), _p(out)))
    return out.reshape(N1, N1, N1)


def gradfft(a, L1, dim):
    N1 = a.shape[0]
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty(N1 ** 3)
    _chk(lib().ref_gradfft(N1, L1, _p(a.ravel()), _p(out), dim))
    return out.reshape(a.shape)


def measure_spectrum(a, L1, n_bin):
    """measure_spectrum (field_statistics.cpp:20-90) -> (kmode[n_bin], power[n_bin])."""
    N1 = a.shape[0]
    a = np.ascontiguousarray(a, dtype=np.float64)
    kmode, power = np.empty(n_bin), np.empty(n_bin)
    _chk(lib().ref_measure_spectrum(N1, L1, _p(a.ravel()), n_bin, _p(kmode), _p(power)))
    return kmode, power


def gradfindif(a, L1, dim):
    N1 = a.shape[0]
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty(N1 ** 3)
    _chk(lib().ref_gradfindif(N1, L1, _p(a.ravel()), _p(out), dim))
    return out.reshape(a.shape)


def fft_backend() -> str:
    return lib().ref_fft_backend().decode()


def num_threads() -> int:
    return int(lib().ref_num_threads())


def fft_stats(reset: bool = False):
    """(wall seconds inside fftw_execute, number of calls) since the last reset -- the FFT shim's own counters,
    so that a timing of the reference can say how much of it is the substitute FFT."""
    fn = lib().shim_fftw_stats
    fn.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int]
    fn.restype = None
    sec, calls = C.c_double(), C.c_long()
    fn(C.byref(sec), C.byref(calls), 1 if reset else 0)
    return sec.value, calls.value

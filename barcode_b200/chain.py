"""`Chain`: one HMC chain's device state behind the C ABI (one `bgpu_handle`).

Thin Python view of include/barcode_gpu.h used by the tests and bench.py; the
method names are the reference's (HMC.cc / HMC_momenta.cc / HMC_mass.cc).
Arrays are numpy float64, shape (N1, N2, N3) or flat, z fastest.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

import numpy as np

from . import _lib
from ._lib import BGPU_CALC_H_EXACT, BgpuError, BgpuParams  # noqa: F401


@dataclass
class Params:
    """input.par / DATA fields the path reads; defaults = bgpu_default_params()."""
    N1: int = 64
    L1: float = 200.0
    xllc: float = 0.0
    yllc: float = 0.0
    zllc: float = 0.0
    xobs: float = 90.0
    yobs: float = 90.0
    zobs: float = 90.0
    planepar: bool = True
    periodic: bool = True
    masskernel: int = 1
    likelihood: int = 1
    sfmodel: int = 1
    rsd_model: bool = False
    calc_h: int = 0
    mass_type: int = 1
    D1: float = 1.0
    D2: float = -3.0 / 7.0 * 0.272 ** (-1.0 / 143.0)   # init_par.cc:526-528 at z = 0
    slength: float = 4.0
    particle_kernel_h_rel: float = 1.0
    ascale: float = 1.0
    OM: float = 0.272
    OL: float = 0.728
    rho_c: float = 1.0
    biasP: float = 1.0
    biasE: float = 1.0
    deltaQ_factor: float = 1.0
    correct_delta: bool = True
    mass_factor: float = 1.0
    div_dH_by_N: bool = False
    delta_min: float = -0.999
    N_bin: int = 200
    device: int = 0

    def to_c(self) -> BgpuParams:
        p = BgpuParams()
        _lib.load().bgpu_default_params(C.byref(p))
        for f in fields(self):
            if f.name in ("N1", "L1"):
                continue
            setattr(p, f.name, type(getattr(p, f.name))(getattr(self, f.name)))
        p.N1 = p.N2 = p.N3 = int(self.N1)
        p.L1 = p.L2 = p.L3 = float(self.L1)
        return p


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} elements, got {a.size}")
    return a


class Chain:
    def __init__(self, params: Params):
        self.params = params
        self.L = _lib.load()
        self.N1 = int(params.N1)
        self.N = self.N1 ** 3
        self.Nhalf = self.N1 * self.N1 * (self.N1 // 2 + 1)
        self._h = C.c_void_p()
        cp = params.to_c()
        _lib.check(self.L.bgpu_create(C.byref(cp), C.byref(self._h)))

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if self._h:
            self.L.bgpu_destroy(self._h)
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

    @property
    def kshape(self):
        """Shape of a k-space (half-complex) array: [x][y][z <= N/2]."""
        return (self.N1, self.N1, self.N1 // 2 + 1)

    # -- inputs -----------------------------------------------------------
    def set_static(self, Power=None, nobs=None, noise=None, window=None):
        arrs = [None if a is None else _f64(a, self.N) for a in (Power, nobs, noise, window)]
        _lib.check(self.L.bgpu_set_static(self._h, *[None if a is None else _dp(a) for a in arrs]))

    def set_mass(self, mass_f=None, mass_r=None):
        arrs = [None if a is None else _f64(a, self.N) for a in (mass_f, mass_r)]
        _lib.check(self.L.bgpu_set_mass(self._h, *[None if a is None else _dp(a) for a in arrs]))

    def hamiltonian_mass(self, signal=None):
        """Hamiltonian_mass (HMC_mass.cc:315-368): returns (mass_f, mass_r).  `signal` (hd->x) is needed by the
        likelihood-force masses, mass_type 2 / 3."""
        mf, mr = np.zeros(self.N), np.zeros(self.N)
        if signal is None:
            _lib.check(self.L.bgpu_hamiltonian_mass(self._h, _dp(mf), _dp(mr)))
        else:
            _lib.check(self.L.bgpu_hamiltonian_mass_x(self._h, _dp(_f64(signal, self.N)), _dp(mf), _dp(mr)))
        return mf.reshape(self.shape), mr.reshape(self.shape)

    # -- the seams S1..S5 -------------------------------------------------
    def gradient_psi(self, signal):
        s = _f64(signal, self.N)
        out = np.empty(self.N)
        _lib.check(self.L.bgpu_gradient_psi(self._h, _dp(s), _dp(out)))
        return out.reshape(self.shape)

    def psi(self, signal, want_deltaX=True):
        s = _f64(signal, self.N)
        a, b = C.c_double(), C.c_double()
        dX = np.empty(self.N) if want_deltaX else None
        _lib.check(self.L.bgpu_psi(self._h, _dp(s), C.byref(a), C.byref(b), _dp(dX) if want_deltaX else None))
        return a.value, b.value, (dX.reshape(self.shape) if want_deltaX else None)

    def kinetic_term(self, momenta):
        p = _f64(momenta, self.N)
        k = C.c_double()
        _lib.check(self.L.bgpu_kinetic(self._h, _dp(p), C.byref(k)))
        return k.value

    def leapfrog(self, s_i, p_i, Neps: int, epsilon: float):
        s, p = _f64(s_i, self.N), _f64(p_i, self.N)
        sf, pf = np.empty(self.N), np.empty(self.N)
        _lib.check(self.L.bgpu_leapfrog(self._h, _dp(s), _dp(p), int(Neps), float(epsilon), _dp(sf), _dp(pf)))
        return sf.reshape(self.shape), pf.reshape(self.shape)

    def delta_hamiltonian(self, s_i, p_i, s_f, p_f):
        """delta_Hamiltonian (HMC.cc:209-248), host arithmetic on the six device-computed scalars."""
        Ki = self.kinetic_term(p_i)
        pri, lki, _ = self.psi(s_i, want_deltaX=False)
        Kf = self.kinetic_term(p_f)
        prf, lkf, dX = self.psi(s_f, want_deltaX=True)
        Hi = Ki + (pri + lki)
        Hf = Kf + (prf + lkf)
        dH = Hf - Hi
        if self.params.div_dH_by_N:
            dH /= float(self.N)
        return dH, dict(dK=Kf - Ki, dE=(prf + lkf) - (pri + lki), dprior=prf - pri, dlikeli=lkf - lki,
                        psi_prior_i=pri, psi_prior_f=prf, psi_likeli_i=lki, psi_likeli_f=lkf,
                        H_kin_i=Ki, H_kin_f=Kf), dX

    def color_momenta(self, white=None, real_gauss=None):
        w = None if white is None else np.ascontiguousarray(white, dtype=np.complex128).reshape(-1)
        g = None if real_gauss is None else _f64(real_gauss, self.N)
        n_full = int(self.params.N1) ** 3   # a slab rank passes the full grid too (barcode_gpu.h)
        if w is not None and w.size != n_full:
            raise ValueError("white noise must be the full N1^3 complex grid")
        out = np.empty(self.N)
        _lib.check(self.L.bgpu_color_momenta(
            self._h, None if w is None else w.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)),
            None if g is None else _dp(g), _dp(out)))
        return out.reshape(self.shape)

    def draw_momenta_device(self, seed: int, draw_index: int):
        """Momenta ~ N(0, M) from the device generator (Philox4x32-10); not GSL-seed-compatible."""
        out = np.empty(self.N)
        _lib.check(self.L.bgpu_draw_momenta_device(self._h, int(seed), int(draw_index), _dp(out)))
        return out.reshape(self.shape)

    def draw_momenta_device_dev(self, seed: int, draw_index: int, d_momenta_ptr: int):
        _lib.check(self.L.bgpu_draw_momenta_device_dev(self._h, int(seed), int(draw_index), C.c_void_p(d_momenta_ptr)))

    # -- device-resident HMC candidate (HamiltonianMC's loop body, HMC.cc:436-506) ---------------------------
    def set_signal(self, x):
        _lib.check(self.L.bgpu_set_signal(self._h, _dp(_f64(x, self.N))))

    def candidate(self, seed: int, draw_index: int, Neps: int, epsilon: float):
        """Momenta from the device generator, Neps leapfrog steps from the current signal, the six energies of
        delta_Hamiltonian -> dict; nothing but scalars crosses PCIe."""
        E = (C.c_double * 6)()
        pf0 = C.c_double()
        _lib.check(self.L.bgpu_candidate(self._h, int(seed), int(draw_index), int(Neps), float(epsilon), E,
                                         C.byref(pf0)))
        keys = ("H_kin_i", "psi_prior_i", "psi_likeli_i", "H_kin_f", "psi_prior_f", "psi_likeli_f")
        out = dict(zip(keys, (float(v) for v in E)))
        out["dH"] = (out["H_kin_f"] + (out["psi_prior_f"] + out["psi_likeli_f"])) - \
            (out["H_kin_i"] + (out["psi_prior_i"] + out["psi_likeli_i"]))
        out["momenta_f0"] = pf0.value
        return out

    def accept(self):
        """Make the candidate the current signal; returns (x, deltaX)."""
        x, dX = np.empty(self.N), np.empty(self.N)
        _lib.check(self.L.bgpu_accept(self._h, _dp(x), _dp(dX)))
        return x.reshape(self.shape), dX.reshape(self.shape)

    def device_normals(self, seed: int, draw_index: int, stream: int, first: int, n: int):
        out = np.empty(n)
        _lib.check(self.L.bgpu_device_normals(self._h, int(seed), int(draw_index), int(stream), int(first), int(n),
                                              _dp(out)))
        return out

    def measure_spectrum(self, signal, n_bin: int):
        """measure_spectrum (field_statistics.cpp:20-90) -> (kmode[n_bin], power[n_bin])."""
        s = _f64(signal, self.N)
        kmode, power = np.empty(n_bin), np.empty(n_bin)
        _lib.check(self.L.bgpu_measure_spectrum(self._h, _dp(s), int(n_bin), _dp(kmode), _dp(power)))
        return kmode, power

    def mock_data(self, seed: int, window_type: int = 1, data_model: int = 0, sigma_min: float = 1.0,
                  sigma_fac: float = 0.0, negative_obs: bool = False):
        """setup_random_test (barcoderunner.cc:42-205) on the device: returns dict(delta_lag, delta_eul, nobs, noise,
        window) and leaves nobs / noise / window as the chain's static inputs."""
        class MP(C.Structure):
            _fields_ = [("window_type", C.c_int), ("data_model", C.c_int), ("sigma_min", C.c_double),
                        ("sigma_fac", C.c_double), ("negative_obs", C.c_int)]
        mp = MP(int(window_type), int(data_model), float(sigma_min), float(sigma_fac), int(bool(negative_obs)))
        out = {k: np.empty(self.N) for k in ("delta_lag", "delta_eul", "nobs", "noise", "window")}
        _lib.check(self.L.bgpu_mock_data(self._h, int(seed), C.byref(mp), _dp(out["delta_lag"]), _dp(out["delta_eul"]),
                                         _dp(out["nobs"]), _dp(out["noise"]), _dp(out["window"])))
        return {k: v.reshape(self.shape) for k, v in out.items()}

    def initial_guess(self, seed: int, kind: int = 2, smoothing_scale: float = 0.0):
        """make_initial_guess (barcoderunner.cc:207-247) on the device: kind 0 zeros, 2 GRF, 3 smoothed GRF, 4 noise."""
        s = np.empty(self.N)
        _lib.check(self.L.bgpu_initial_guess(self._h, int(seed), int(kind), float(smoothing_scale), _dp(s)))
        return s.reshape(self.shape)

    def forward(self, signal, want_pos=False):
        s = _f64(signal, self.N)
        dX = np.empty(self.N)
        if want_pos:
            x, y, z = np.empty(self.N), np.empty(self.N), np.empty(self.N)
            _lib.check(self.L.bgpu_forward(self._h, _dp(s), _dp(dX), _dp(x), _dp(y), _dp(z)))
            return dX.reshape(self.shape), x, y, z
        _lib.check(self.L.bgpu_forward(self._h, _dp(s), _dp(dX), None, None, None))
        return dX.reshape(self.shape)

    # -- building blocks for parity tests ---------------------------------
    def assign_density(self, x, y, z):
        x, y, z = _f64(x, self.N), _f64(y, self.N), _f64(z, self.N)
        rho = np.empty(self.N)
        _lib.check(self.L.bgpu_assign_density(self._h, _dp(x), _dp(y), _dp(z), _dp(rho)))
        return rho.reshape(self.shape)

    def cell_indices(self, x, y, z):
        x, y, z = _f64(x), _f64(y), _f64(z)
        n = x.size
        ci, cj, ck = (np.empty(n, dtype=np.int32) for _ in range(3))
        ip = C.POINTER(C.c_int)
        _lib.check(self.L.bgpu_cell_indices(self._h, _dp(x), _dp(y), _dp(z), n, ci.ctypes.data_as(ip),
                                            cj.ctypes.data_as(ip), ck.ctypes.data_as(ip)))
        return ci, cj, ck

    def fft_r2c(self, a):
        a = _f64(a, self.N)
        out = np.empty(2 * self.Nhalf)
        _lib.check(self.L.bgpu_fft_r2c(self._h, _dp(a), _dp(out)))
        return out.view(np.complex128).reshape(self.kshape)

    def fft_c2r(self, c):
        c = np.ascontiguousarray(c, dtype=np.complex128).reshape(-1)
        if c.size != self.Nhalf:
            raise ValueError("expected the half-complex array")
        out = np.empty(self.N)
        _lib.check(self.L.bgpu_fft_c2r(self._h, c.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), _dp(out)))
        return out.reshape(self.shape)

    def convolve_inv_corr(self, signal, corr):
        s, c = _f64(signal, self.N), _f64(corr, self.N)
        out = np.empty(self.N)
        _lib.check(self.L.bgpu_convolve_inv_corr(self._h, _dp(s), _dp(c), _dp(out)))
        return out.reshape(self.shape)

    # -- device-pointer variants (torch tensors on this chain's device) -----
    def set_stream(self, cuda_stream_ptr: int):
        _lib.check(self.L.bgpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        _lib.check(self.L.bgpu_synchronize(self._h))

    def gradient_psi_dev(self, d_signal_ptr: int, d_grad_ptr: int):
        _lib.check(self.L.bgpu_gradient_psi_dev(self._h, C.c_void_p(d_signal_ptr), C.c_void_p(d_grad_ptr)))

    def leapfrog_dev(self, d_signal_ptr: int, d_momenta_ptr: int, Neps: int, epsilon: float):
        _lib.check(self.L.bgpu_leapfrog_dev(self._h, C.c_void_p(d_signal_ptr), C.c_void_p(d_momenta_ptr), int(Neps),
                                            float(epsilon)))

    def psi_dev(self, d_signal_ptr: int, d_deltaX_ptr: int = 0):
        a, b = C.c_double(), C.c_double()
        _lib.check(self.L.bgpu_psi_dev(self._h, C.c_void_p(d_signal_ptr), C.byref(a), C.byref(b),
                                       C.c_void_p(d_deltaX_ptr) if d_deltaX_ptr else None))
        return a.value, b.value

    def kinetic_dev(self, d_momenta_ptr: int):
        k = C.c_double()
        _lib.check(self.L.bgpu_kinetic_dev(self._h, C.c_void_p(d_momenta_ptr), C.byref(k)))
        return k.value


def kernel_launches() -> int:
    return int(_lib.load().bgpu_kernel_launches())


def profile_begin():
    """Start recording CUDA events around every kernel-class launch (roofline leg of bench.py)."""
    _lib.check(_lib.load().bgpu_profile_begin())


def profile_end():
    """Stop recording; returns {kernel class: (total ms, launches)}."""
    L = _lib.load()
    n = _lib.PROFILE_KINDS
    ms = (C.c_double * n)()
    cnt = (C.c_uint64 * n)()
    _lib.check(L.bgpu_profile_end(ms, cnt, n))
    return {L.bgpu_profile_kind_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}

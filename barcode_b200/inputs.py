"""Host-side input preparation that sits either side of the GPU path.

These are the cold, once-per-run steps the reference does on the host before the
sampler starts (barcode/main.cc:150-168, barlib/src/calc_power.cc:31-108,
barlib/src/barcoderunner.cc:42-205).  They stay on the host here too; only the
synthetic-data helpers used by bench.py and the tests touch the GPU, and they
do it through the same C ABI as everything else.
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PK_TABLE = os.path.join(_HERE, "data", "pk_table.npz")   # package data (written by tests/golden/make_golden.py)


def calc_ki(N: int, L: float) -> np.ndarray:
    """k_i = 2 pi i / L for i <= N/2, else -2 pi (N - i) / L (scale_space.cpp:41-51)."""
    kfac = 2.0 * np.pi / L
    i = np.arange(N)
    return np.where(i <= N // 2, kfac * i, -kfac * (N - i))


def load_pk_table(path: str = PK_TABLE):
    """(k, P) of the tabulated linear power spectrum, as float32 -- the precision the
    reference reads its CAMB table in (calc_power.cc:41-42)."""
    with np.load(path) as f:
        return f["k"].astype(np.float32), f["P"].astype(np.float32)


def power_on_grid(k_tab, p_tab, N: int, L: float, chunk: int = 32) -> np.ndarray:
    """readtab (calc_power.cc:31-108): linear interpolation of the table at |k| on the
    full N^3 real-indexed grid, P(0) = 0.  Chunked over x so 512^3 fits in host RAM."""
    kt = np.asarray(k_tab, dtype=np.float32).astype(np.float64)
    pt = np.asarray(p_tab, dtype=np.float32).astype(np.float64)
    k = calc_ki(N, L)
    out = np.empty((N, N, N))
    for i0 in range(0, N, chunk):
        i1 = min(N, i0 + chunk)
        # k_squared (scale_space.cpp:16-38) sums kx^2 + ky^2 + kz^2 left to right
        k2 = (k[i0:i1, None, None] ** 2 + k[None, :, None] ** 2) + k[None, None, :] ** 2
        ktot = np.sqrt(k2).ravel()
        idx = np.clip(np.searchsorted(kt, ktot, side="right") - 1, 0, len(kt) - 2)
        x_lo, x_hi = kt[idx], kt[idx + 1]
        y_lo, y_hi = pt[idx], pt[idx + 1]
        out[i0:i1] = (y_lo + (ktot - x_lo) / (x_hi - x_lo) * (y_hi - y_lo)).reshape(i1 - i0, N, N)
    out[0, 0, 0] = 0.0
    return out


def power_on_planes(k_tab, p_tab, N: int, L: float, x0: int, nx: int) -> np.ndarray:
    """The planes [x0, x0 + nx) of power_on_grid: what one rank of a slab-decomposed chain holds."""
    kt = np.asarray(k_tab, dtype=np.float32).astype(np.float64)
    pt = np.asarray(p_tab, dtype=np.float32).astype(np.float64)
    k = calc_ki(N, L)
    out = np.empty((nx, N, N))
    for c0 in range(0, nx, 16):
        c1 = min(nx, c0 + 16)
        k2 = (k[x0 + c0:x0 + c1, None, None] ** 2 + k[None, :, None] ** 2) + k[None, None, :] ** 2
        ktot = np.sqrt(k2).ravel()
        idx = np.clip(np.searchsorted(kt, ktot, side="right") - 1, 0, len(kt) - 2)
        x_lo, x_hi = kt[idx], kt[idx + 1]
        y_lo, y_hi = pt[idx], pt[idx + 1]
        out[c0:c1] = (y_lo + (ktot - x_lo) / (x_hi - x_lo) * (y_hi - y_lo)).reshape(c1 - c0, N, N)
    if x0 == 0:
        out[0, 0, 0] = 0.0
    return out


def power_on_kslab(k_tab, p_tab, N: int, L: float, y0: int, ny: int) -> np.ndarray:
    """P(|k|) on the transposed k-space slab [x][y0 <= y < y0 + ny][z <= N/2] (slab.SlabChain.kshape)."""
    kt = np.asarray(k_tab, dtype=np.float32).astype(np.float64)
    pt = np.asarray(p_tab, dtype=np.float32).astype(np.float64)
    k = calc_ki(N, L)
    nzh = N // 2 + 1
    out = np.empty((N, ny, nzh))
    for c0 in range(0, N, 16):
        c1 = min(N, c0 + 16)
        k2 = (k[c0:c1, None, None] ** 2 + k[None, y0:y0 + ny, None] ** 2) + k[None, None, :nzh] ** 2
        ktot = np.sqrt(k2).ravel()
        idx = np.clip(np.searchsorted(kt, ktot, side="right") - 1, 0, len(kt) - 2)
        x_lo, x_hi = kt[idx], kt[idx + 1]
        y_lo, y_hi = pt[idx], pt[idx + 1]
        out[c0:c1] = (y_lo + (ktot - x_lo) / (x_hi - x_lo) * (y_hi - y_lo)).reshape(c1 - c0, ny, nzh)
    if y0 == 0:
        out[0, 0, 0] = 0.0
    return out


def slab_problem(sc, seed: int = 1, signal_scale: float = 0.5):
    """`synthetic_problem` for a slab-decomposed chain, every array a local slab: Gaussian random
    fields are made with the chain's own distributed FFT (white noise -> r2c -> * sqrt(P n / V) ->
    c2r), nobs = max(0, 1 + forward(truth) + N(0, 1)), window = noise = 1."""
    N, L = sc.N1, sc.params.L1
    k_tab, p_tab = load_pk_table()
    P = power_on_planes(k_tab, p_tab, N, L, sc.x0, sc.Ns)
    Pk = power_on_kslab(k_tab, p_tab, N, L, sc.x0, sc.Ns)
    amp = np.sqrt(np.maximum(Pk, 0.0) * float(N) ** 3 / L ** 3)
    ones = np.ones(sc.shape)
    sc.set_static(Power=P, nobs=ones, noise=ones, window=ones)

    def grf(sd):
        w = np.random.default_rng(1000003 * sd + sc.rank).standard_normal(sc.shape)
        return sc.fft_c2r(sc.fft_r2c(w) * amp)

    truth = grf(seed)
    d_eul = sc.forward(truth)
    nobs = np.maximum(0.0, 1.0 + d_eul + np.random.default_rng(7 * seed + 1000 + sc.rank).standard_normal(sc.shape))
    sc.set_static(nobs=nobs)
    signal = signal_scale * grf(seed + 1)
    sc.hamiltonian_mass()
    return dict(Power=P, nobs=nobs, signal=signal)


def box_length(N: int) -> float:
    """The shipped cell size: 200 Mpc/h over 64 cells (data/input.par:123-125)."""
    return N * (200.0 / 64.0)


def complex_white_noise(N: int, seed: int) -> np.ndarray:
    """Unit complex Gaussian white noise on the full N^3 grid (production mode: NOT the
    GSL mt19937 shell-ordered stream of random.hpp:36-120; for seed parity feed the
    reference's stream to Chain.color_momenta instead)."""
    rng = np.random.default_rng(seed)
    out = np.empty((N, N, N), dtype=np.complex128)
    v = out.view(np.float64)
    for i in range(N):
        v[i] = rng.standard_normal((N, 2 * N))
    return out


def synthetic_problem(chain, seed: int = 1, signal_scale: float = 0.5):
    """The reference's `random_test` recipe (barcoderunner.cc:42-205) on the GPU path:
    truth = GRF(Power); delta_eul = forward(truth); window = 1; Gaussian: sigma = 1,
    nobs = max(0, 1 + delta_eul + N(0,1)); Poisson: nobs ~ Poisson(1 + delta_eul);
    mass = Hamiltonian_mass.  Returns dict(Power, nobs, noise, window, signal, momenta):
    evaluation signal = an independent GRF draw times `signal_scale`, momenta from the
    mass.  White noise comes from numpy (see complex_white_noise)."""
    p = chain.params
    N, L = p.N1, p.L1
    k_tab, p_tab = load_pk_table()
    power = power_on_grid(k_tab, p_tab, N, L)
    ones = np.ones(N ** 3)
    chain.set_static(Power=power, nobs=ones, noise=ones, window=ones)
    # GRF with spectrum Power: colour through a temporary "mass == P" view of the same kernel
    saved_mass_type = p.mass_type
    truth = _garfield(chain, power, seed)
    d_eul = chain.forward(truth)
    rng = np.random.default_rng(seed + 1000)
    if p.likelihood == 1:
        nobs = np.maximum(0.0, 1.0 + d_eul + rng.standard_normal(d_eul.shape))
    else:
        nobs = rng.poisson(np.maximum(1.0 + d_eul, 0.0)).astype(np.float64)
    chain.set_static(nobs=nobs, noise=ones, window=ones)
    signal = signal_scale * _garfield(chain, power, seed + 1)
    chain.hamiltonian_mass()
    gauss = np.random.default_rng(seed + 3000).standard_normal(N ** 3) if p.mass_type == 0 else None
    white = complex_white_noise(N, seed + 2) if p.mass_type != 0 else None
    momenta = chain.color_momenta(white, gauss)
    assert saved_mass_type == p.mass_type
    return dict(Power=power, nobs=nobs, noise=ones.reshape(d_eul.shape), window=ones.reshape(d_eul.shape),
                signal=signal, momenta=momenta, truth=truth)


def _garfield(chain, spectrum, seed: int):
    """create_GARFIELD(spectrum) through the chain's colouring kernel: load `spectrum` as
    the Fourier mass, colour, and restore the real mass afterwards is the caller's job
    (synthetic_problem calls hamiltonian_mass() after its last draw)."""
    from .chain import Chain, Params
    p = chain.params
    if p.mass_type in (1, 4):
        chain.set_mass(mass_f=spectrum)
        return chain.color_momenta(complex_white_noise(p.N1, seed), None)
    # real-space-mass chains cannot colour: use a throw-away Fourier-mass chain
    q = Params(**{**p.__dict__, "mass_type": 4})
    with Chain(q) as tmp:
        tmp.set_mass(mass_f=spectrum)
        return tmp.color_momenta(complex_white_noise(p.N1, seed), None)

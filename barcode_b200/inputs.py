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
PK_TABLE = os.path.join(os.path.dirname(_HERE), "tests", "golden", "pk_table.npz")


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

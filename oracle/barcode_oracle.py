"""oracle/barcode_oracle.py -- TEST INFRASTRUCTURE, not product code.

CPU (numpy) restatement of the algorithm on Barcode's HMC gradient-and-leapfrog
hot path.  Every function cites the reference lines it follows (paths relative
to /root/reference/).  It is pinned against the compiled reference itself
(``oracle/_ref``, built by ``oracle/Makefile`` from the unmodified sources) in
``tests/test_oracle_ref.py`` and against the golden vectors that library
produced (``tests/golden/``, generator ``tests/golden/make_golden.py``).  The
reference's own test-suite holds NO golden vector or known-answer test for this
path, so those two are the pins.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
leg may import this module; the product path (``barcode_b200``) never does.

Conventions: arrays are ``float64[N1,N2,N3]``, C order
(z fastest, ``idx = k + N3*(j + N2*i)``, disp_part.cc:60); DFT is
``FOURIER_DEF_2``: forward unnormalised, inverse times 1/N (fftwrapper.cc:100-101).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EPS_K2 = 1.0e-14  # define_opt.h:82


# --------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------
@dataclass
class Params:
    """The subset of DATA / HAMIL_DATA the path reads (struct_hamil.h:51-212)."""
    N1: int
    L1: float
    masskernel: int = 1        # 0 NGP, 1 CIC, 2 TSC (massFunctions.cc)
    likelihood: int = 1        # 0 Poisson, 1 Gaussian, 2 log-normal, 3 Gaussian random field (init_par.cc:534-559)
    rsd_model: bool = False
    calc_h: int = 0            # 0, 1, 2 (SPH adjoint), 3 (its Fourier / TSC variant): reference; 4 = exact NGP/CIC/TSC adjoint (new)
    mass_type: int = 1         # 0 ones (R), 1 1/P (FS), 4 P (FS)  (HMC_mass.cc:315-368)
    D1: float = 1.0
    D2: float = -3.0 / 7.0 * 0.272 ** (-1.0 / 143.0)  # init_par.cc:526-528 at z = 0: -3/7 D1^2 Omega^(-1/143)
    sfmodel: int = 1           # 1 Zel'dovich; anything else -> Lag2Eul_non_zeldovich (Lag2Eul.cc:329-331)
    slength: float = 4.0       # n->kth = data->numerical->slength (struct_hamil.h:259): ALPT smoothing radius
    particle_kernel_h_rel: float = 1.0   # SPH scale length in cells (init_par.cc:379; masskernel 3)
    ascale: float = 1.0
    OM: float = 0.272          # init_par.cc:38,482 (cmbcosm = 3)
    OL: float = 0.728
    deltaQ_factor: float = 1.0
    correct_delta: bool = True
    mass_factor: float = 1.0
    min1: float = 0.0
    min2: float = 0.0
    min3: float = 0.0
    rho_c: float = 1.0         # init_par.cc:574-578
    biasP: float = 1.0
    biasE: float = 1.0
    div_dH_by_N: bool = False
    delta_min: float = -0.999  # log-normal density floor (data/input.par:51)
    N_bin: int = 200           # measure_spectrum bins (data/input.par:129)

    @property
    def N(self):
        return self.N1 ** 3

    @property
    def d(self):
        return self.L1 / float(self.N1)

    @property
    def vol(self):
        return self.L1 * self.L1 * self.L1

    @property
    def kernel_h(self):  # init_par.cc:377-379 (cubic cells)
        return self.particle_kernel_h_rel * self.d

    @property
    def mass_fs(self):  # struct_hamil.h:276-313
        return self.mass_type in (1, 2, 3, 4, 5)

    @property
    def mass_rs(self):
        return self.mass_type in (0, 5, 6, 60)


# --------------------------------------------------------------------------
# k-space helpers
# --------------------------------------------------------------------------
def calc_ki(N: int, L: float) -> np.ndarray:
    """scale_space.cpp:41-51 -- k_i = 2 pi i / L (i <= N/2) else -2 pi (N - i) / L."""
    kfac = 2.0 * np.pi / L
    i = np.arange(N)
    return np.where(i <= N // 2, kfac * i, -kfac * (N - i))


def k_grids(N: int, L: float):
    """(kx, ky, kz_half) broadcastable to [N, N, N/2+1]."""
    k = calc_ki(N, L)
    return k[:, None, None], k[None, :, None], k[None, None, : N // 2 + 1]


def k_squared_full(N: int, L: float) -> np.ndarray:
    """scale_space.cpp:16-38 on the full real-indexed grid."""
    k = calc_ki(N, L)
    return k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, :] ** 2


def nyquist_mask(N: int) -> np.ndarray:
    """i==N1/2 || j==N2/2 || k==N3/2 on the half grid (EqSolvers.cc:254, gradient.cpp:69,205)."""
    i = np.arange(N)
    kh = np.arange(N // 2 + 1)
    return (i[:, None, None] == N // 2) | (i[None, :, None] == N // 2) | (kh[None, None, :] == N // 2)


def rfft(a):
    return np.fft.rfftn(a)


def irfft(c, N):
    """fftC2R: inverse including 1/N (fftwrapper.cc:26-53)."""
    return np.fft.irfftn(c, s=(N, N, N), axes=(0, 1, 2))


def power_on_grid(k_tab, p_tab, N: int, L: float) -> np.ndarray:
    """calc_power.cc:31-108 (readtab): the table is read into *float* arrays
    (:41-42), promoted to double, linearly interpolated (gsl_interp_linear) at
    |k| on the full grid; P(0) = 0 (:102-103)."""
    kt = np.asarray(k_tab, dtype=np.float32).astype(np.float64)
    pt = np.asarray(p_tab, dtype=np.float32).astype(np.float64)
    ktot = np.sqrt(k_squared_full(N, L))
    idx = np.clip(np.searchsorted(kt, ktot.ravel(), side="right") - 1, 0, len(kt) - 2)
    x_lo, x_hi = kt[idx], kt[idx + 1]
    y_lo, y_hi = pt[idx], pt[idx + 1]
    out = y_lo + (ktot.ravel() - x_lo) / (x_hi - x_lo) * (y_hi - y_lo)
    out = out.reshape(N, N, N)
    out[0, 0, 0] = 0.0
    return out


def read_power_table(fname: str):
    tab = np.loadtxt(fname)
    return tab[:, 0], tab[:, 1]


# --------------------------------------------------------------------------
# A5: inverse-correlation convolution (prior gradient, mass apply)
# --------------------------------------------------------------------------
def convolve_inv_corr(p: Params, signal, corr) -> np.ndarray:
    """HMC_help.cc:16-64 -- IFFT[(V/N)/C(k) FFT[signal]], C <= 0 -> 0; the
    correlation array is indexed as the REAL grid at (i, j, k <= N3/2); no
    Nyquist zeroing."""
    N = p.N1
    normFS = p.vol / float(p.N)
    c = corr.reshape(N, N, N)[:, :, : N // 2 + 1]
    fac = np.where(c > 0.0, normFS / np.where(c > 0.0, c, 1.0), 0.0)
    return irfft(rfft(signal.reshape(N, N, N)) * fac, N)


def grad_log_prior(p: Params, signal, power) -> np.ndarray:
    """hmc/prior/gaussian.cpp:15-18."""
    return convolve_inv_corr(p, signal, power)


def log_prior(p: Params, signal, power) -> float:
    """hmc/prior/gaussian.cpp:20-35 -- 1/2 sum s * S^-1 s."""
    return float(np.sum(0.5 * signal.reshape(-1) * convolve_inv_corr(p, signal, power).reshape(-1)))


# --------------------------------------------------------------------------
# A7/A8: Zel'dovich displacement
# --------------------------------------------------------------------------
def theta2vel(p: Params, f):
    """EqSolvers.cc:168-277 -- Psi_c = IFFT[(k_c/k^2) (Im f^, -Re f^)], zero at
    k^2 <= 1e-14 and on every Nyquist plane; norm=false so cpecvel = 1."""
    N = p.N1
    fh = rfft(f.reshape(N, N, N))
    kx, ky, kz = k_grids(N, p.L1)
    ksq = kx * kx + ky * ky + kz * kz
    ok = (ksq > EPS_K2) & ~nyquist_mask(N)
    fac = np.where(ok, 1.0 / np.where(ok, ksq, 1.0), 0.0)
    rot = fh.imag - 1j * fh.real  # (Im, -Re) = -i f^
    out = []
    for kc in (kx, ky, kz):
        out.append(irfft((fac * kc) * rot, N))
    return out


def pacman(x, L):
    """pacman.cpp:20-28."""
    x = np.array(x, dtype=np.float64, copy=True)
    neg = x < 0.0
    x[neg] = np.fmod(x[neg], L)
    x[neg] += L
    big = x >= L
    x[big] = np.fmod(x[big], L)
    return x


def E_hubble(a, OM, OL):
    """E_Hubble_a, cosmo.cc:26-31: sqrt(OM/(a*a*a) + OK/(a*a) + OL)."""
    OK = 1.0 - OM - OL
    return np.sqrt(OM / (a * a * a) + OK / (a * a) + OL)


def fgrow(a, OM, OL):
    """cosmo.cc:182-217, term 1: f = Omega(a)^(5/9)."""
    E = E_hubble(a, OM, OL)
    Omega = OM / ((E * E) * (a * a * a))
    return Omega ** (5.0 / 9.0)


def c_pecvel(a, OM, OL):
    """cosmo.cc:220-235: f * 100 * E * a."""
    return fgrow(a, OM, OL) * 100.0 * E_hubble(a, OM, OL) * a


def positions(p: Params, psi):
    """disp_part.cc:55-126 (reggrid): x = d*i + 0.5*d, += Psi, pacman; then
    rsd.cc:18-67 (plane-parallel, periodic): z += (cpecvel*Psi_z) * (1/Hub/a), pacman."""
    N, d, L = p.N1, p.d, p.L1
    g = d * np.arange(N, dtype=np.float64) + 0.5 * d
    x = pacman(g[:, None, None] + psi[0], L)
    y = pacman(g[None, :, None] + psi[1], L)
    z = pacman(g[None, None, :] + psi[2], L)
    if p.rsd_model:
        # Lag2Eul.cc:378-381 + rsd.cc:27-28,39,52-57
        vez = c_pecvel(p.ascale, p.OM, p.OL) * psi[2]
        OC = 1.0 - p.OM - p.OL
        Hub = 100.0 * np.sqrt(p.OM / p.ascale / p.ascale / p.ascale + p.OL + OC / p.ascale / p.ascale)
        v_norm = 1.0 / Hub / p.ascale
        z = pacman(z + vez * v_norm, L)
    return x, y, z


# --------------------------------------------------------------------------
# A12: mass assignment
# --------------------------------------------------------------------------
def _in_domain(p: Params, x, y, z, closed_upper: bool):
    L = p.L1
    if closed_upper:  # TSC: massFunctions.cc:195
        return ((x >= p.min1) & (x <= p.min1 + L) & (y >= p.min2) & (y <= p.min2 + L)
                & (z >= p.min3) & (z <= p.min3 + L))
    return ((x >= p.min1) & (x < p.min1 + L) & (y >= p.min2) & (y < p.min2 + L)
            & (z >= p.min3) & (z < p.min3 + L))


def cic_cells_weights(p: Params, x):
    """interpolate_grid.cpp:27-79 for one coordinate: cell (i, i+), weights (tx, dx)."""
    N, d, L = p.N1, p.d, p.L1
    xpos = pacman(x - 0.5 * d, L)
    i = (xpos / d).astype(np.uint64)          # static_cast<ULONG>
    i = (i + np.uint64(N)) % np.uint64(N)
    ip = (i + np.uint64(1)) % np.uint64(N)
    dx = xpos / d - i.astype(np.float64)
    tx = 1.0 - dx
    return i.astype(np.int64), ip.astype(np.int64), tx, dx


def ngp_cells(p: Params, x, xmin):
    """massFunctions.cc:72-79: floor((x-min)/d) -> unsigned -> fmod(i, N)."""
    N, d = p.N1, p.d
    i = np.floor((x - xmin) / d).astype(np.int64).astype(np.uint32)
    return np.fmod(i.astype(np.float64), float(N)).astype(np.int64)


def tsc_cells_weights(p: Params, x, xmin):
    """massFunctions.cc:198-235: cells (i-1, i, i+1), weights (hm1, h0, hp1)."""
    N, d = p.N1, p.d
    i = ngp_cells(p, x, xmin)
    ip = np.fmod((i + 1).astype(np.float64), float(N)).astype(np.int64)
    im = np.fmod((i - 1 + N).astype(np.float64), float(N)).astype(np.int64)
    xc = i.astype(np.float64) + 0.5
    dx = (x - xmin) / d - xc
    h0 = 0.75 - dx * dx
    hp = 0.5 * (0.5 + dx) * (0.5 + dx)
    hm = 0.5 * (0.5 - dx) * (0.5 - dx)
    return (im, i, ip), (hm, h0, hp), dx


def density(p: Params, x, y, z) -> np.ndarray:
    """getDensity_NGP / _CIC / _TSC (massFunctions.cc:49-98, 100-164, 167-364),
    unit particle masses.  Summation order differs from the OpenMP atomics of
    the reference (which is itself run-to-run nondeterministic, main.cc:87-89)."""
    N = p.N1
    x, y, z = (np.asarray(a, dtype=np.float64).ravel() for a in (x, y, z))
    rho = np.zeros(N * N * N)
    if p.masskernel == 0:
        ok = _in_domain(p, x, y, z, False)
        i, j, k = ngp_cells(p, x[ok], p.min1), ngp_cells(p, y[ok], p.min2), ngp_cells(p, z[ok], p.min3)
        np.add.at(rho, k + N * (j + N * i), 1.0)
    elif p.masskernel == 1:
        ok = _in_domain(p, x, y, z, False)
        i0, i1, tx, dx = cic_cells_weights(p, x[ok])
        j0, j1, ty, dy = cic_cells_weights(p, y[ok])
        k0, k1, tz, dz = cic_cells_weights(p, z[ok])
        for ii, wx in ((i0, tx), (i1, dx)):
            for jj, wy in ((j0, ty), (j1, dy)):
                for kk, wz in ((k0, tz), (k1, dz)):
                    # mass*tx*ty*tz evaluated left to right (massFunctions.cc:129-157)
                    np.add.at(rho, kk + N * (jj + N * ii), (1.0 * wx) * wy * wz)
    elif p.masskernel == 2:
        ok = _in_domain(p, x, y, z, True)
        ci, wi, _ = tsc_cells_weights(p, x[ok], p.min1)
        cj, wj, _ = tsc_cells_weights(p, y[ok], p.min2)
        ck, wk, _ = tsc_cells_weights(p, z[ok], p.min3)
        for a in range(3):
            for b in range(3):
                for c in range(3):
                    np.add.at(rho, ck[c] + N * (cj[b] + N * ci[a]), (1.0 * wi[a]) * wj[b] * wk[c])
    elif p.masskernel == 3:
        rho = density_sph(p, x, y, z).ravel()
    else:
        raise NotImplementedError("masskernel %d" % p.masskernel)
    return rho.reshape(N, N, N)


def sph_kernel(r, h):
    """SPH_kernel_3D (massFunctions.cc:366-384): Monaghan W_4 spline."""
    q = r / h
    a = 1.0 / np.pi / (h * h * h)
    return np.where(q <= 1.0, a * (1 - 3.0 / 2 * q * q + 3.0 / 4 * q * q * q),
                    np.where(q <= 2.0, a * (1.0 / 4 * (2.0 - q) ** 3), 0.0))


def density_sph(p: Params, x, y, z):
    """getDensity_SPH (massFunctions.cc:392-495): every particle adds W(r, h) to the cells within
    reach = int(2h/d)+1 of its own cell whose centre is within 2h; unit masses, no normalisation."""
    N, d, h = p.N1, p.d, p.kernel_h
    ok = _in_domain(p, x, y, z, False)
    x, y, z = x[ok], y[ok], z[ok]
    reach = int(2 * h / d) + 1
    ix, iy, iz = (x / d).astype(np.int64), (y / d).astype(np.int64), (z / d).astype(np.int64)
    ccx, ccy, ccz = (ix + 0.5) * d, (iy + 0.5) * d, (iz + 0.5) * d
    rho = np.zeros(N * N * N)
    for i1 in range(-reach, reach + 1):
        dx = x - (ccx + i1 * d)
        kx = (N + i1 + ix) % N
        for i2 in range(-reach, reach + 1):
            dy = y - (ccy + i2 * d)
            ky = (N + i2 + iy) % N
            for i3 in range(-reach, reach + 1):
                dz = z - (ccz + i3 * d)
                kz = (N + i3 + iz) % N
                r = np.sqrt(dx * dx + dy * dy + dz * dz)
                m = r / h <= 2.0
                np.add.at(rho, (kz + N * (ky + N * kx))[m], sph_kernel(r[m], h))
    return rho.reshape(N, N, N)


def sph_hull(p: Params):
    """SPH_kernel_3D_cells + SPH_kernel_3D_cells_hull_1 (SPH_kernel.cpp:62-139): the (i, j) columns
    and their inclusive k range that can hold a cell centre within 2h of any point of the central cell."""
    d, h = p.d, p.kernel_h
    reach = int(2 * h / d) + 1
    hull = {}
    for i1 in range(-reach, reach + 1):
        for i2 in range(-reach, reach + 1):
            for i3 in range(-reach, reach + 1):
                r_sq = ((abs(i1) - 0.5) * d) ** 2 + ((abs(i2) - 0.5) * d) ** 2 + ((abs(i3) - 0.5) * d) ** 2
                if r_sq <= (2 * h) ** 2:
                    lo, hi = hull.get((i1, i2), (i3, i3))
                    hull[(i1, i2)] = (min(lo, i3), max(hi, i3))
    return hull


def calc_V_sph(p: Params, r, x, y, z):
    """likelihood_calc_V_SPH (HMC_models.cc:200-303): V_p = (rho_c V/N) sum_cells r_c gradW((x_p - x_c)/h)
    with gradW = partial * (x_p - x_c)/h / (pi h^4) (grad_SPH_kernel_3D_h_units, SPH_kernel.cpp:148-208);
    z component times (1 + f) under plane-parallel RSD."""
    N, d, h = p.N1, p.d, p.kernel_h
    x, y, z = (np.asarray(a, dtype=np.float64).ravel() for a in (x, y, z))
    r = np.asarray(r, dtype=np.float64).reshape(N, N, N)
    norm = 1.0 / (np.pi * h ** 4)
    ix, iy, iz = (x / d).astype(np.int64), (y / d).astype(np.int64), (z / d).astype(np.int64)
    h_inv = 1.0 / h
    d_h = d * h_inv
    dpc = [x * h_inv - (ix + 0.5) * d_h, y * h_inv - (iy + 0.5) * d_h, z * h_inv - (iz + 0.5) * d_h]
    V = [np.zeros_like(x) for _ in range(3)]
    for (i1, i2), (k0, k1) in sph_hull(p).items():
        dxh = dpc[0] - i1 * d_h
        dyh = dpc[1] - i2 * d_h
        kx, ky = (ix + i1) % N, (iy + i2) % N
        for i3 in range(k0, k1 + 1):
            dzh = dpc[2] - i3 * d_h
            kz = (iz + i3) % N
            q_sq = dxh * dxh + dyh * dyh + dzh * dzh
            q = np.sqrt(q_sq)
            with np.errstate(divide="ignore", invalid="ignore"):
                partial = np.where(q_sq > 4, 0.0, np.where(q_sq > 1, -0.75 * (q - 2) * (q - 2) * norm / q,
                                                           (2.25 * q - 3) * norm))
            c = r[kx, ky, kz] * partial
            V[0] += c * dxh
            V[1] += c * dyh
            V[2] += c * dzh
    normalize = p.rho_c * p.vol / float(p.N)
    V = [normalize * v for v in V]
    if p.rsd_model:
        V[2] = V[2] + fgrow(p.ascale, p.OM, p.OL) * V[2]
    shp = (N, N, N)
    return [v.reshape(shp) for v in V]


def sph_kernel_fourier(p: Params) -> np.ndarray:
    """h * SPH_kernel_F on the half grid, as likelihood_calc_V_SPH_fourier_TSC writes it
    (HMC_models_testing.cpp:62-105): norm = 24/h^3 * rho_c V/N, the kernel transform
    (3 + cos 2k - k sin k + cos k (k sin k - 4)) / k^6 in the PHYSICAL k (not k h), 1/h^3 at k = 0.
    The numerator cancels to ~k^6/240, so at the grid's small k its value is rounding noise that depends on the C
    library's last bits: evaluated here with the same C library calls (math.sin / cos / sqrt), operation by
    operation, so that the restatement and the compiled reference agree on one machine."""
    import math
    N, h = p.N1, p.kernel_h
    norm = (24.0 / (h * h * h)) * (p.rho_c * p.vol / float(p.N))
    k1 = calc_ki(N, p.L1)
    out = np.empty((N, N, N // 2 + 1))
    for i in range(N):
        kx = float(k1[i])
        for j in range(N):
            ky = float(k1[j])
            for k in range(N // 2 + 1):
                kz = float(k1[k])
                k_sq = kx * kx + ky * ky + kz * kz
                if k_sq == 0.0:
                    f = 1.0 / (h * h * h)
                else:
                    kk = math.sqrt(k_sq)
                    ksink = kk * math.sin(kk)
                    f = norm * (3 + math.cos(2 * kk) - ksink + math.cos(kk) * (ksink - 4)) / (k_sq * k_sq * k_sq)
                out[i, j, k] = f
    return out


def interpolate_tsc_buggy(p: Params, x, y, z, field) -> np.ndarray:
    """interpolate_TSC (interpolate_grid.cpp:134-190), bug-compatible: the upper weights of x and y are computed
    from dz (:166-167).  Cell = (unsigned)(x/d), weights 3/4 - D^2 and 1/2 (3/2 - |D -+ 1|)^2."""
    N, d = p.N1, p.d
    x, y, z = (np.asarray(a, dtype=np.float64).ravel() for a in (x, y, z))
    f = np.asarray(field, dtype=np.float64).reshape(N, N, N)
    xk, yk, zk = x / d, y / d, z / d
    ix, iy, iz = xk.astype(np.int64), yk.astype(np.int64), zk.astype(np.int64)
    dx, dy, dz = xk - (ix + 0.5), yk - (iy + 0.5), zk - (iz + 0.5)
    wx = [0.5 * (1.5 - np.abs(dx + 1)) ** 2, 0.75 - dx * dx, 0.5 * (1.5 - np.abs(dz - 1)) ** 2]
    wy = [0.5 * (1.5 - np.abs(dy + 1)) ** 2, 0.75 - dy * dy, 0.5 * (1.5 - np.abs(dz - 1)) ** 2]
    wz = [0.5 * (1.5 - np.abs(dz + 1)) ** 2, 0.75 - dz * dz, 0.5 * (1.5 - np.abs(dz - 1)) ** 2]
    cx = [(ix - 1 + N) % N, ix % N, (ix + 1) % N]
    cy = [(iy - 1 + N) % N, iy % N, (iy + 1) % N]
    cz = [(iz - 1 + N) % N, iz % N, (iz + 1) % N]
    out = np.zeros_like(x)
    for a in range(3):
        for b in range(3):
            for c in range(3):
                out += wx[a] * wy[b] * wz[c] * f[cx[a], cy[b], cz[c]]
    return out


def calc_V_sph_fourier_tsc(p: Params, r, x, y, z):
    """likelihood_calc_V_SPH_fourier_TSC (HMC_models_testing.cpp:54-188; calc_h = 3): the residual convolved with
    the gradient of the SPH kernel in k-space -- i k_c h W^(k) r^(k), no Nyquist zeroing -- and TSC-interpolated to
    the particle positions; z component times (1 + f) under plane-parallel RSD."""
    N = p.N1
    kx, ky, kz = k_grids(N, p.L1)
    rh = rfft(np.asarray(r, dtype=np.float64).reshape(N, N, N))
    hW = p.kernel_h * sph_kernel_fourier(p)
    V = []
    for kc in (kx, ky, kz):
        conv = irfft((kc * hW) * (-rh.imag + 1j * rh.real), N)
        V.append(interpolate_tsc_buggy(p, x, y, z, conv))
    if p.rsd_model:
        V[2] = V[2] + fgrow(p.ascale, p.OM, p.OL) * V[2]
    return [v.reshape(N, N, N) for v in V]


def overdens(rho) -> np.ndarray:
    """massFunctions.cc:30-47: rho / mean - 1, mean accumulated in double."""
    mean = float(np.sum(rho, dtype=np.float64)) / float(rho.size)
    return rho / mean - 1.0


def poisson_solver(p: Params, delta):
    """PoissonSolver (EqSolvers.cc:29-64): IFFT[-1/k^2 FFT[delta]], 0 at k^2 == 0; NO Nyquist zeroing."""
    N = p.N1
    dh = rfft(delta.reshape(N, N, N))
    kx, ky, kz = k_grids(N, p.L1)
    ksq = kx * kx + ky * ky + kz * kz
    fac = np.where(ksq > 0, -1.0 / np.where(ksq > 0, ksq, 1.0), 0.0)
    return irfft(dh * fac, N)


def calc_m2v(p: Params, phi):
    """calc_m2v_mem with GFINDIFF (EqSolvers.cc:373-422): second derivatives by applying the 4th-order
    finite difference twice, then the sum of the three 2x2 minors of the Hessian."""
    dx = gradfindif(p, phi, 1)
    Lxx, Lxy, Lxz = gradfindif(p, dx, 1), gradfindif(p, dx, 2), gradfindif(p, dx, 3)
    dy = gradfindif(p, phi, 2)
    Lyy, Lyz = gradfindif(p, dy, 2), gradfindif(p, dy, 3)
    Lzz = gradfindif(p, gradfindif(p, phi, 3), 3)
    return Lxx * Lyy - Lxy * Lxy + Lxx * Lzz - Lxz * Lxz + Lyy * Lzz - Lyz * Lyz


def alpt_kernel(p: Params):
    """kernelcomp with filtertype 1 (convolution.cpp:224-322): exp(-k^2 rS^2 / 2) on the grid, divided by
    the sum of its inverse transform (= its k = 0 value, 1, up to rounding)."""
    N = p.N1
    kx, ky, kz = k_grids(N, p.L1)
    return np.exp(-(kx * kx + ky * ky + kz * kz) * (p.slength * p.slength) / 2.0)


def convcomp(p: Params, a, kern):
    """convcomp (convolution.cpp:327-377): IFFT[kernel * FFT[a]] (the reference does it with full c2c
    transforms and the kernel re-read from `auxkernelr<int(slength)>.dat`)."""
    N = p.N1
    return irfft(rfft(a.reshape(N, N, N)) * kern, N)


def displacement_non_zeldovich(p: Params, s):
    """Lag2Eul_non_zeldovich up to Psi (Lag2Eul.cc:138-268): theta = D1 s - D2 delta2 smoothed by K,
    spherical-collapse divergence on the small scales, then cellboundcomp (massFunctions.cc:588-660):
    every Psi averaged with its (i-1, j-1, k-1) neighbour, periodically."""
    N = p.N1
    s = s.reshape(N, N, N)
    phi1 = poisson_solver(p, s)
    d2 = calc_m2v(p, phi1)
    kern = alpt_kernel(p)
    lpt = convcomp(p, p.D1 * s - p.D2 * d2, kern)
    psilin = -p.D1 * s
    arg = 1.0 + 2.0 / 3.0 * psilin
    sc = -np.where(arg > 0.0, 3.0 * (np.sqrt(np.where(arg > 0.0, arg, 0.0)) - 1.0), -3.0)
    A = theta2vel(p, lpt)
    B = theta2vel(p, sc)
    out = []
    for a, b in zip(A, B):
        tot = (a + b) - convcomp(p, b, kern)
        out.append(0.5 * (np.roll(tot, (1, 1, 1), (0, 1, 2)) + tot))
    return out


def forward(p: Params, signal):
    """likelihood_grad_log_like's forward half (HMC_models.cc:383-406) ->
    Lag2Eul_zeldovich / _rsd_zeldovich (Lag2Eul.cc:69-132, 338-424).
    Returns (deltaX, (x, y, z), (Psi_x, Psi_y, Psi_z))."""
    N = p.N1
    s = signal.reshape(N, N, N)
    if p.deltaQ_factor != 1.0:
        s = p.deltaQ_factor * s
    if p.sfmodel == 1 or p.rsd_model:   # the reference runs Zel'dovich under rsd_model for any sfmodel
        psi = theta2vel(p, -p.D1 * s)
    else:
        psi = displacement_non_zeldovich(p, s)
    x, y, z = positions(p, psi)
    rho = density(p, x, y, z)
    return overdens(rho), (x, y, z), psi


# --------------------------------------------------------------------------
# A13: likelihoods
# --------------------------------------------------------------------------
def partial_f(p: Params, deltaX, nobs, noise, window, exact_sign: bool = False) -> np.ndarray:
    """Residual per cell.  Gaussian: gaussian_independent.cpp:24-42,
    r = (n - Lambda)/sigma^2 where w > 0 and Lambda > 0.  Poisson:
    poissonian.cpp:19-34, r = (1 - n/Lambda) rho_c bE bP dens^(bE-1) where w > 0
    and dens > 0.  `exact_sign` flips the Poisson residual to the Gaussian
    convention r = -d(-lnL)/d delta (SURVEY A.4 sign trap) for the new exact
    adjoint; it is never used in the reference-parity modes."""
    dX = deltaX.reshape(-1)
    n, sg, w = nobs.reshape(-1), noise.reshape(-1), window.reshape(-1)
    out = np.zeros_like(dX)
    if p.likelihood == 1:
        Lam = w * p.rho_c * np.power(1.0 + p.biasP * dX, p.biasE)
        ok = (w > 0.0) & (Lam > 0.0)
        out[ok] = (n[ok] - Lam[ok]) / (sg[ok] * sg[ok])
    elif p.likelihood == 0:
        dens = 1.0 + p.biasP * dX
        ok = (w > 0.0) & (dens > 0.0)
        Lam = w[ok] * p.rho_c * np.power(dens[ok], p.biasE)
        out[ok] = (1 - n[ok] / Lam) * p.rho_c * p.biasE * p.biasP * np.power(dens[ok], p.biasE - 1)
        if exact_sign:
            out = -out
    elif p.likelihood == 2:
        # lognormal_independent.cpp:41-55: r = (n - Lambda)/sigma^2 where w > 0, Lambda = log(rho_c (1 + bP
        # delta)^bE) of the UNCLAMPED density (nan / inf where it is not positive, as in the reference)
        ok = w > 0.0
        with np.errstate(invalid="ignore", divide="ignore"):
            Lam = np.log(p.rho_c * np.power(1.0 + p.biasP * dX[ok], p.biasE))
        out[ok] = (n[ok] - Lam) / (sg[ok] * sg[ok])
        if exact_sign:  # NEW: -d(-lnL)/d delta of the clamped value the reference's log_like sums (:112-122)
            dc = np.maximum(dX[ok], p.delta_min)
            Lv = np.log(p.rho_c * (1.0 + dc))
            out[ok] = np.where(dX[ok] < p.delta_min, 0.0, (n[ok] - Lv) / (sg[ok] * sg[ok]) / (1.0 + dX[ok]))
    else:
        raise NotImplementedError("likelihood %d" % p.likelihood)
    return out.reshape(deltaX.shape)


def lognormal_f(p: Params, deltaX) -> np.ndarray:
    """lognormal_likelihood_f_delta_x (lognormal_independent.cpp:57-79): log(rho_c (1 + max(delta, delta_min)))."""
    return np.log(p.rho_c * (1.0 + np.maximum(deltaX, p.delta_min)))


def neg_log_like_from_delta(p: Params, deltaX, nobs, noise, window) -> float:
    """gaussian_independent.cpp:80-91 / poissonian.cpp:60-73."""
    dX = deltaX.reshape(-1)
    n, sg, w = nobs.reshape(-1), noise.reshape(-1), window.reshape(-1)
    if p.likelihood == 1:
        Lam = w * p.rho_c * np.power(1.0 + p.biasP * dX, p.biasE)
        ok = (w > 0.0) & (Lam > 0.0)
        return float(np.sum(0.5 * ((Lam[ok] - n[ok]) / sg[ok]) ** 2))
    if p.likelihood == 0:
        dens = 1.0 + p.biasP * dX
        with np.errstate(invalid="ignore"):
            Lam = w * p.rho_c * np.power(dens, p.biasE)
        ok = (w > 0.0) & (Lam > 0.0)
        return float(np.sum(Lam[ok] - n[ok] * np.log(Lam[ok])))
    if p.likelihood == 2:  # lognormal_independent.cpp:112-122
        ok = w > 0.0
        res = lognormal_f(p, dX[ok]) - n[ok]
        return float(np.sum(0.5 * res * res / (sg[ok] * sg[ok])))
    raise NotImplementedError


def log_like(p: Params, signal, nobs, noise, window):
    """gaussian_likelihood_log_like (gaussian_independent.cpp:51-92) /
    poissonian_likelihood_log_like (poissonian.cpp:44-74).  The Poisson path
    never applies deltaQ_factor or RSD (poissonian.cpp:54-56).  Returns
    (-lnL, deltaX)."""
    if p.likelihood == 3:  # gaussian_random_field.cpp:40-52: the Lagrangian field itself, no forward model
        s_, n, sg, w = signal.reshape(-1), nobs.reshape(-1), noise.reshape(-1), window.reshape(-1)
        ok = w > 0.0
        return float(np.sum(0.5 * ((s_[ok] - n[ok]) / sg[ok]) ** 2)), None
    if p.likelihood in (0, 2):  # neither applies deltaQ_factor or RSD (poissonian.cpp:54-56, lognormal_independent.cpp:98-110)
        q = Params(**{**p.__dict__, "rsd_model": False, "deltaQ_factor": 1.0})
        dX, _, _ = forward(q, signal)
    else:
        dX, _, _ = forward(p, signal)
    return neg_log_like_from_delta(p, dX, nobs, noise, window), dX


# --------------------------------------------------------------------------
# A14: adjoint chain
# --------------------------------------------------------------------------
def gradfft(p: Params, a, dim: int) -> np.ndarray:
    """gradient.cpp:22-78: IFFT[i k_dim a^], Nyquist planes zeroed."""
    N = p.N1
    ah = rfft(a.reshape(N, N, N))
    kc = k_grids(N, p.L1)[dim - 1]
    out = (1j * kc) * ah
    out = np.where(nyquist_mask(N), 0.0, out)
    return irfft(out, N)


def gradfindif(p: Params, a, dim: int) -> np.ndarray:
    """gradient.cpp:81-153: -fac*((4/3)(f[-1]-f[+1]) - (1/6)(f[-2]-f[+2])), fac = N/(2L)."""
    N = p.N1
    fac = N / (2.0 * p.L1)
    a = a.reshape(N, N, N)
    ax = dim - 1
    l, r = np.roll(a, 1, ax), np.roll(a, -1, ax)
    ll, rr = np.roll(a, 2, ax), np.roll(a, -2, ax)
    return -(fac * ((4.0 / 3) * (l - r) - (1.0 / 6) * (ll - rr)))


def grad_inv_lap_sum(p: Params, V) -> np.ndarray:
    """sum_c IFFT[(k_c/k^2)(Im V^_c, -Re V^_c)] with Nyquist planes zeroed and
    1/k^2 -> 0 at k^2 == 0 (gradient.cpp:157-211, HMC_models.cc:350-371)."""
    N = p.N1
    kx, ky, kz = k_grids(N, p.L1)
    ksq = kx * kx + ky * ky + kz * kz
    fac = np.where(ksq > 0, 1.0 / np.where(ksq > 0, ksq, 1.0), 0.0)
    fac = np.where(nyquist_mask(N), 0.0, fac)
    acc = np.zeros((N, N, N // 2 + 1), dtype=np.complex128)
    for kc, Vc in zip((kx, ky, kz), V):
        vh = rfft(Vc.reshape(N, N, N))
        acc += (kc * fac) * (vh.imag - 1j * vh.real)
    return irfft(acc, N)


def calc_h0(p: Params, deltaX, nobs, noise, window) -> np.ndarray:
    """likelihood_calc_h (HMC_models_testing.cpp:25-50): h = sum_c IFFT[-i k_c/k^2
    FFT[r * d_c deltaX]]; d_c = gradfft (Gaussian, gaussian_independent.cpp:44-50)
    or gradfindif (Poisson, poissonian.cpp:37-42)."""
    r = partial_f(p, deltaX, nobs, noise, window)
    g = gradfft if p.likelihood == 1 else gradfindif
    f = lognormal_f(p, deltaX) if p.likelihood == 2 else deltaX  # lognormal_independent.cpp:81-91
    return grad_inv_lap_sum(p, [r * g(p, f, c) for c in (1, 2, 3)])


def gather_adjoint(p: Params, r, x, y, z):
    """NEW (no reference code): V_p = sum_cells r_c dW_c(x_p)/dx_p for
    the NGP/CIC/TSC assignment weights above; z component times (1+f) under RSD
    (the reference does the same for its SPH adjoint, HMC_models.cc:295-301)."""
    N, d = p.N1, p.d
    r = r.reshape(-1)
    shp = x.shape
    x, y, z = x.ravel(), y.ravel(), z.ravel()
    Vx, Vy, Vz = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    if p.masskernel == 1:
        i0, i1, tx, dx = cic_cells_weights(p, x)
        j0, j1, ty, dy = cic_cells_weights(p, y)
        k0, k1, tz, dz = cic_cells_weights(p, z)
        cx, cy, cz = ((i0, tx, -1.0 / d), (i1, dx, 1.0 / d)), ((j0, ty, -1.0 / d), (j1, dy, 1.0 / d)), \
            ((k0, tz, -1.0 / d), (k1, dz, 1.0 / d))
        for ii, wx, gx in cx:
            for jj, wy, gy in cy:
                for kk, wz, gz in cz:
                    rc = r[kk + N * (jj + N * ii)]
                    Vx += rc * gx * wy * wz
                    Vy += rc * wx * gy * wz
                    Vz += rc * wx * wy * gz
    elif p.masskernel == 2:
        ci, wi, ddx = tsc_cells_weights(p, x, p.min1)
        cj, wj, ddy = tsc_cells_weights(p, y, p.min2)
        ck, wk, ddz = tsc_cells_weights(p, z, p.min3)
        # d/dx of (hm, h0, hp) with Delta = x/d - (i+1/2): (-(1/2-D), -2D, (1/2+D)) / d
        gi = (-(0.5 - ddx) / d, -2.0 * ddx / d, (0.5 + ddx) / d)
        gj = (-(0.5 - ddy) / d, -2.0 * ddy / d, (0.5 + ddy) / d)
        gk = (-(0.5 - ddz) / d, -2.0 * ddz / d, (0.5 + ddz) / d)
        for a in range(3):
            for b in range(3):
                for c in range(3):
                    rc = r[ck[c] + N * (cj[b] + N * ci[a])]
                    Vx += rc * gi[a] * wj[b] * wk[c]
                    Vy += rc * wi[a] * gj[b] * wk[c]
                    Vz += rc * wi[a] * wj[b] * gk[c]
    elif p.masskernel == 0:
        pass  # NGP weights are piecewise constant: derivative is identically zero
    else:
        raise NotImplementedError
    if p.rsd_model:
        Vz = Vz + fgrow(p.ascale, p.OM, p.OL) * Vz
    return Vx.reshape(shp), Vy.reshape(shp), Vz.reshape(shp)


def grad_log_like(p: Params, signal, nobs, noise, window):
    """likelihood_grad_log_like (HMC_models.cc:377-471).  Returns (grad, deltaX)."""
    if p.likelihood == 3:  # HMC.cc:159-160 -> grf_likelihood_grad_log_like (gaussian_random_field.cpp:25-38)
        N = p.N1
        s3, w, sg = signal.reshape(N, N, N), window.reshape(N, N, N), noise.reshape(N, N, N)
        return np.where(w > 0.0, (s3 - nobs.reshape(N, N, N)) / (sg * sg), 0.0), None
    dX, (x, y, z), _ = forward(p, signal)
    if p.calc_h == 0:
        h = calc_h0(p, dX, nobs, noise, window)
    elif p.calc_h == 1:
        h = partial_f(p, dX, nobs, noise, window)
    elif p.calc_h == 2:
        # likelihood_calc_h_SPH (HMC_models.cc:312-372): the reference's exact adjoint, SPH kernel only
        r = partial_f(p, dX, nobs, noise, window)
        h = grad_inv_lap_sum(p, calc_V_sph(p, r, x, y, z))
    elif p.calc_h == 3:
        # the Fourier / TSC variant of the same adjoint (HMC_models.cc:336-340)
        r = partial_f(p, dX, nobs, noise, window)
        h = grad_inv_lap_sum(p, calc_V_sph_fourier_tsc(p, r, x, y, z))
    elif p.calc_h == 4:
        r = partial_f(p, dX, nobs, noise, window, exact_sign=True)
        mean = 1.0  # rho/mean: mean == 1 analytically for NGP/CIC/TSC
        V = gather_adjoint(p, r / mean, x, y, z)
        if not (p.sfmodel == 1 or p.rsd_model):
            return p.deltaQ_factor * non_zeldovich_adjoint(p, p.deltaQ_factor * signal, V), dX
        h = grad_inv_lap_sum(p, V)
    else:
        raise NotImplementedError("calc_h %d" % p.calc_h)
    norm = -1.0 * p.deltaQ_factor          # HMC_models.cc:460-465
    if p.correct_delta:
        norm *= p.D1                       # :467-469
    return norm * h, dX


def non_zeldovich_adjoint(p: Params, s_in, V) -> np.ndarray:
    """NEW (the reference has no adjoint of Lag2Eul_non_zeldovich, HMC_models.cc:458): the exact transpose of
    displacement_non_zeldovich applied to the particle forces V_c = -d(-lnL)/dx_c.  With P_c = IFFT T_c FFT the
    projection (T_c = -i k_c/k^2, P_c^T = -P_c), CB the cell-boundary average and C = K o theta_LPT + (1 - K) o
    theta_SC:   d(-lnL)/d in = (dC/d in)^T sum_c P_c CB^T V_c
      theta_SC' = D1 / sqrt(1 - 2/3 D1 in) where the root's argument is positive, else 0
      theta_LPT = D1 in - D2 m2v(phi), phi = IFFT[-in^/k^2]:  (d theta_LPT/d in)^T u = D1 u - D2 Poisson[G(u)],
      G(u) = sum_ab FD_a FD_b (dm2v/dL_ab u)   (the 4th-order stencil is antisymmetric, so (FD_b FD_a)^T = FD_a FD_b)
    `s_in` = deltaQ_factor * signal, the model's input."""
    N = p.N1
    s_in = s_in.reshape(N, N, N)
    W = [0.5 * (np.roll(v.reshape(N, N, N), (-1, -1, -1), (0, 1, 2)) + v.reshape(N, N, N)) for v in V]   # CB^T
    hW = grad_inv_lap_sum(p, W)
    kern = alpt_kernel(p)
    Ah = rfft(hW)
    u_lpt = irfft(kern * Ah, N)
    u_sc = irfft((1.0 - kern) * Ah, N)
    arg = 1.0 - 2.0 / 3.0 * p.D1 * s_in
    dsc = np.where(arg > 0.0, p.D1 / np.sqrt(np.where(arg > 0.0, arg, 1.0)), 0.0)
    phi = poisson_solver(p, s_in)
    dx, dy, dz = gradfindif(p, phi, 1), gradfindif(p, phi, 2), gradfindif(p, phi, 3)
    Lxx, Lxy, Lxz = gradfindif(p, dx, 1), gradfindif(p, dx, 2), gradfindif(p, dx, 3)
    Lyy, Lyz, Lzz = gradfindif(p, dy, 2), gradfindif(p, dy, 3), gradfindif(p, dz, 3)

    def dd(a, b, f):
        return gradfindif(p, gradfindif(p, f, a), b)

    G = (dd(1, 1, (Lyy + Lzz) * u_lpt) + dd(2, 2, (Lxx + Lzz) * u_lpt) + dd(3, 3, (Lxx + Lyy) * u_lpt)
         - 2.0 * dd(1, 2, Lxy * u_lpt) - 2.0 * dd(1, 3, Lxz * u_lpt) - 2.0 * dd(2, 3, Lyz * u_lpt))
    return p.D1 * u_lpt - p.D2 * poisson_solver(p, G) + dsc * u_sc


def gradient_psi(p: Params, signal, power, nobs, noise, window) -> np.ndarray:
    """HMC.cc:146-206 at the default test knobs (data/input.par:150-158)."""
    return grad_log_prior(p, signal, power) + grad_log_like(p, signal, nobs, noise, window)[0]


def psi(p: Params, signal, power, nobs, noise, window):
    """HMC.cc:124-143: (psi_prior, psi_likelihood, deltaX side effect)."""
    pl, dX = log_like(p, signal, nobs, noise, window)
    return log_prior(p, signal, power), pl, dX


# --------------------------------------------------------------------------
# A16/A17: mass, kinetic energy, leapfrog
# --------------------------------------------------------------------------
def hamiltonian_mass(p: Params, power, signal=None, nobs=None, noise=None, window=None):
    """HMC_mass.cc:315-368 types 0/1/4, and 2/3 (the likelihood-force masses, :39-160, which need the signal and
    the observations): (mass_f, mass_r)."""
    N = p.N1
    mass_f = np.zeros((N, N, N))
    mass_r = np.zeros((N, N, N))
    if p.mass_type == 0:
        mass_r[:] = 1.0
    elif p.mass_type == 1:
        P = power.reshape(N, N, N)
        mass_f = np.where(P > 0.0, 1.0 / np.where(P > 0.0, P, 1.0), 0.0)  # inv_ps, HMC_mass.cc
    elif p.mass_type == 4:
        mass_f = power.reshape(N, N, N).copy()
    elif p.mass_type in (2, 3):
        force = grad_log_like(p, signal.reshape(N, N, N), nobs, noise, window)[0]     # likeli_force_power, :39-51
        kmode, fspec = measure_spectrum(p, force, p.N_bin)
        P = power.reshape(N, N, N)
        invP = np.where(P > 0.0, 1.0 / np.where(P > 0.0, P, 1.0), 0.0)
        k = calc_ki(N, p.L1)
        dk = np.sqrt(3.0 * k[N // 2] * k[N // 2]) / float(p.N_bin)
        if p.mass_type == 2:   # Hamiltonian_mass_likeli_force, :54-83
            kr = np.sqrt((k[:, None, None] ** 2 + k[None, :, None] ** 2) + k[None, None, :] ** 2)
            b = (kr / dk).astype(np.int64)
            # the corner mode's bin index equals N_bin: the reference reads one past its array there; 0 here
            F = np.where((kr > 0.0) & (b < p.N_bin), fspec[np.minimum(b, p.N_bin - 1)], 0.0)
        else:                  # Hamiltonian_mass_mean_likeli_force, :86-114
            F = np.sum(4.0 * np.pi * kmode * kmode * dk * fspec) / np.sum(4.0 * np.pi * kmode * kmode * dk)
        mass_f = 2.0 * invP + np.sqrt(invP * F)
    else:
        raise NotImplementedError("mass_type %d is off the hot path (SURVEY section 2 row 5)" % p.mass_type)
    if p.mass_fs:
        mass_f = p.mass_factor * mass_f
    return mass_f, mass_r


def apply_inv_mass(p: Params, mom, mass_f, mass_r) -> np.ndarray:
    """HMC.cc:298-327 / :69-99: M^-1 p."""
    N = p.N1
    out = convolve_inv_corr(p, mom, mass_f) if p.mass_fs else np.zeros((N, N, N))
    if p.mass_rs:
        mr = mass_r.reshape(N, N, N)
        out = out + np.where(mr > 0.0, 1.0 / np.where(mr > 0.0, mr, 1.0), 0.0) * mom.reshape(N, N, N)
    return out


def kinetic_term(p: Params, mom, mass_f, mass_r) -> float:
    """HMC.cc:64-121."""
    return float(np.sum(0.5 * mom.reshape(-1) * apply_inv_mass(p, mom, mass_f, mass_r).reshape(-1)))


def leapfrog(p: Params, s_i, p_i, Neps: int, eps: float, power, nobs, noise, window, mass_f, mass_r):
    """Hamiltonian_EoM (HMC.cc:251-369) after the two RNG draws."""
    N = p.N1
    s = s_i.reshape(N, N, N).copy()
    m = p_i.reshape(N, N, N).copy()
    g = gradient_psi(p, s, power, nobs, noise, window)
    for _ in range(Neps):
        m -= 0.5 * eps * g
        s += eps * apply_inv_mass(p, m, mass_f, mass_r)
        g = gradient_psi(p, s, power, nobs, noise, window)
        m -= 0.5 * eps * g
        if abs(m.flat[0]) > 1e50:   # HMC.cc:360-364
            break
    return s, m


def delta_hamiltonian(p: Params, s_i, p_i, s_f, p_f, power, nobs, noise, window, mass_f, mass_r):
    """HMC.cc:209-248."""
    Ki = kinetic_term(p, p_i, mass_f, mass_r)
    pri, lki, _ = psi(p, s_i, power, nobs, noise, window)
    Kf = kinetic_term(p, p_f, mass_f, mass_r)
    prf, lkf, dX = psi(p, s_f, power, nobs, noise, window)
    dH = (Kf + (prf + lkf)) - (Ki + (pri + lki))
    if p.div_dH_by_N:
        dH /= float(p.N)
    return dH, dict(dK=Kf - Ki, dE=(prf + lkf) - (pri + lki), dprior=prf - pri, dlikeli=lkf - lki,
                    psi_prior_i=pri, psi_prior_f=prf, psi_likeli_i=lki, psi_likeli_f=lkf, H_kin_i=Ki, H_kin_f=Kf), dX


# --------------------------------------------------------------------------
# A2: momentum draw
# --------------------------------------------------------------------------
def gsl_mt19937_gaussians(seed: int, count: int) -> np.ndarray:
    """`count` successive gsl_ran_ugaussian values of gsl_rng_mt19937 seeded with
    `seed`: polar Box-Muller; every attempt consumes exactly two
    uniform_pos draws, so the k-th output is the k-th accepted attempt of the raw
    stream and the whole thing vectorises."""
    from numpy.random import MT19937
    m = int(count / 0.78) + 1024
    while True:
        bg = MT19937()
        bg._legacy_seeding(seed if seed != 0 else 4357)
        raw = bg.random_raw(2 * m).astype(np.float64) / 4294967296.0
        if np.any(raw == 0.0):
            raise RuntimeError("uniform_pos rejection hit (p = 2^-32 per draw): use a scalar loop")
        x = -1 + 2 * raw[0::2]
        y = -1 + 2 * raw[1::2]
        r2 = x * x + y * y
        idx = np.nonzero(~((r2 > 1.0) | (r2 == 0)))[0]
        if len(idx) >= count:
            idx = idx[:count]
            return y[idx] * np.sqrt(-2.0 * np.log(r2[idx]) / r2[idx])
        m *= 2


def white_noise_shell_order(N: int, gauss: np.ndarray) -> np.ndarray:
    """resolution_independent_random_grid_FS<double>(N, rng, half_size=false)
    (random.hpp:36-120): complex white noise on the FULL N^3 grid, filled shell
    by shell so that an N grid is a sub-cube of the 2N grid; each entry takes two
    successive Gaussians (re, im).  Pure-Python loop: small N only."""
    out = np.zeros((N, N, N), dtype=np.complex128)
    it = iter(range(0, 2 * N ** 3, 2))

    def nxt():
        q = next(it)
        return complex(gauss[q], gauss[q + 1])

    g = N
    for i in range(g // 2):
        for k in range(i + 1):
            for j in range(i):            # "slim" side, random.hpp:68-81
                for a, b, c in ((i, j, k), (g - 1 - i, j, k), (i, g - 1 - j, k), (g - 1 - i, g - 1 - j, k),
                                (i, j, g - 1 - k), (g - 1 - i, j, g - 1 - k), (i, g - 1 - j, g - 1 - k),
                                (g - 1 - i, g - 1 - j, g - 1 - k)):
                    out[a, b, c] = nxt()
            for j in range(i + 1):        # "broad" side, :82-95
                for a, b, c in ((j, i, k), (g - 1 - j, i, k), (j, g - 1 - i, k), (g - 1 - j, g - 1 - i, k),
                                (j, i, g - 1 - k), (g - 1 - j, i, g - 1 - k), (j, g - 1 - i, g - 1 - k),
                                (g - 1 - j, g - 1 - i, g - 1 - k)):
                    out[a, b, c] = nxt()
        for j in range(i):                # "roof", :97-111
            for k in range(i):
                for a, b, c in ((j, k, i), (g - 1 - j, k, i), (j, g - 1 - k, i), (g - 1 - j, g - 1 - k, i),
                                (j, k, g - 1 - i), (g - 1 - j, k, g - 1 - i), (j, g - 1 - k, g - 1 - i),
                                (g - 1 - j, g - 1 - k, g - 1 - i)):
                    out[a, b, c] = nxt()
    return out


def colour_half_grid(p: Params, white: np.ndarray, spec: np.ndarray) -> np.ndarray:
    """create_GARFIELD's colouring + Hermitian symmetrisation (random.cpp:102-496)
    restated for the entries that survive into the half array (c <= N/2);
    sigma = sqrt(N^2/V * spec/2) read at the folded index (:81-83,104).

    Rule (each half-array element is written exactly once by the reference's
    27-case loop, so the loop is data parallel):
      * the 8 self-conjugate corners: DC = 0 (:347-351), others Re *= sqrt(2) sigma,
        Im = 0 (:243-247, 439-483);
      * 0 < c < N/2: sigma*W(a,b,c), except a > N/2 and b > N/2 (both interior):
        sigma*conj(W(N-a, N-b, N-c)) (:136-140);
      * c in {0, N/2}: if b interior: b < N/2 keeps sigma*W, b > N/2 takes
        sigma*conj(W(N-a mod N, N-b, c)); if b in {0, N/2} and a interior:
        a < N/2 keeps, a > N/2 takes sigma*conj(W(N-a, b, c)).
    """
    N = p.N1
    h = N // 2
    W = white.reshape(N, N, N)
    S = spec.reshape(N, N, N)
    amp = float(p.N) * float(p.N) / p.vol
    a = np.arange(N)[:, None, None]
    b = np.arange(N)[None, :, None]
    c = np.arange(h + 1)[None, None, :]
    fa, fb = np.where(a > h, N - a, a), np.where(b > h, N - b, b)
    sigma = np.sqrt(amp * S[fa, fb, c] / 2.0)
    a_fix, b_fix, c_fix = (a == 0) | (a == h), (b == 0) | (b == h), (c == 0) | (c == h)
    ma, mb = (N - a) % N, (N - b) % N
    own = W[:, :, : h + 1]
    out = np.empty((N, N, h + 1), dtype=np.complex128)
    # interior c
    mirror3 = W[ma, mb, (N - c) % N]
    use_m3 = (~c_fix) & (a > h) & (b > h) & (~a_fix) & (~b_fix)
    # boundary c planes
    mirror2 = W[ma, mb, c]
    use_m2 = c_fix & (((~b_fix) & (b > h)) | (b_fix & (~a_fix) & (a > h)))
    src = np.where(use_m3, mirror3, np.where(use_m2, mirror2, own))
    conj = use_m3 | use_m2
    re = src.real * sigma
    im = src.imag * sigma
    out.real = re
    out.imag = np.where(conj, -im, im)
    corner = a_fix & b_fix & c_fix
    cre = own.real * (np.sqrt(2.0) * sigma)
    out.real = np.where(corner, cre, out.real)
    out.imag = np.where(corner, 0.0, out.imag)
    out[0, 0, 0] = 0.0
    return out


def create_garfield(p: Params, white: np.ndarray, spec: np.ndarray) -> np.ndarray:
    """random.cpp:48-511 after the RNG: colour, half copy (:498-507), fftC2R."""
    return irfft(colour_half_grid(p, white, spec), p.N1)


def draw_momenta(p: Params, white: np.ndarray, mass_f, mass_r, real_gauss=None) -> np.ndarray:
    """HMC_momenta.cc:42-92: coloured field with spectrum mass_f (+ sqrt(mass_r)*N(0,1))."""
    N = p.N1
    mom = create_garfield(p, white, mass_f) if p.mass_fs else np.zeros((N, N, N))
    if p.mass_rs:
        mom = mom + np.sqrt(mass_r.reshape(N, N, N)) * real_gauss.reshape(N, N, N)
    return mom


# --------------------------------------------------------------------------
# F3: measure_spectrum (field_statistics.cpp:20-90)
# --------------------------------------------------------------------------
def measure_spectrum(p: Params, signal, n_bin: int):
    """Spherically binned power of the FULL complex transform: bin = (ULONG)(|k| / dk), dk = |k(N/2, N/2, N/2)| /
    n_bin, modes at or beyond kmax dropped; kmode = mean |k| of a bin, power = mean |F|^2 * V / N^2
    (FOURIER_DEF_2); empty bins stay 0.  Returns (kmode, power)."""
    N = p.N1
    F = np.fft.fftn(signal.reshape(N, N, N))
    k = calc_ki(N, p.L1)
    ktot = np.sqrt((k[:, None, None] ** 2 + k[None, :, None] ** 2) + k[None, None, :] ** 2)
    kny = k[N // 2]
    dk = np.sqrt(3.0 * kny * kny) / float(n_bin)
    b = (ktot / dk).astype(np.int64).ravel()
    ok = b < n_bin
    nmode = np.bincount(b[ok], minlength=n_bin).astype(np.float64)
    kmode = np.bincount(b[ok], weights=ktot.ravel()[ok], minlength=n_bin)
    power = np.bincount(b[ok], weights=(np.abs(F) ** 2).ravel()[ok], minlength=n_bin)
    have = nmode > 0
    kmode[have] /= nmode[have]
    power[have] = power[have] / nmode[have] * (p.vol / float(N) ** 3 / float(N) ** 3)
    return kmode, power


# --------------------------------------------------------------------------
# F4: counter-based Gaussians of the device momentum draw (no reference code: the reference draws from
# GSL's serial mt19937).  Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw, SC'11; Random123),
# pinned by its known-answer vectors in tests/test_host.py.
# --------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised over the counter words (uint64 arrays holding 32-bit values); returns four uint64 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = _PHILOX_M0 * c0, _PHILOX_M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK32, p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _PHILOX_W0) & 0xFFFFFFFF, (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def device_normals(seed: int, draw: int, stream: int, first: int, n: int) -> np.ndarray:
    """Elements [first, first + n) of draw `draw`, stream `stream` under `seed` (kernels.cu philox_normal_kernel):
    pair g = element // 2 uses counter {g_lo, g_hi, draw_lo, draw_hi ^ stream << 24}, key {seed_lo, seed_hi};
    u = 53-bit uniforms in (0, 1); Box-Muller."""
    g = np.arange(first // 2, (first + n) // 2, dtype=np.uint64)
    z = np.zeros_like(g)
    r = philox4x32_10(g & _MASK32, g >> np.uint64(32), z + np.uint64(draw & 0xFFFFFFFF),
                      z + np.uint64(((draw >> 32) ^ (stream << 24)) & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)

    def u53(hi, lo):
        return ((hi >> np.uint64(5)).astype(np.float64) * 67108864.0 + (lo >> np.uint64(6)).astype(np.float64) + 0.5) \
            / 9007199254740992.0

    u1, u2 = u53(r[0], r[1]), u53(r[2], r[3])
    rad = np.sqrt(-2.0 * np.log(u1))
    out = np.empty(n)
    out[0::2] = rad * np.cos(2.0 * np.pi * u2)
    out[1::2] = rad * np.sin(2.0 * np.pi * u2)
    return out

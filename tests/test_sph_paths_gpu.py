"""The SPH spline kernel (masskernel 3) and its exact adjoint (calc_h = 2) on every code path of the GPU library,
against the numpy restatement of getDensity_SPH / likelihood_calc_V_SPH (oracle/barcode_oracle.py, itself pinned to the
compiled reference by tests/test_oracle_ref.py and to the golden vectors za_sph_*):

  h / d in [0.75, 1.25)   column list + unrolled z loop   (particles_sph.cu *_cols5_kernel; the shipped default h = d)
  wider hulls             column list, z loop of any half-range   (*_cols_kernel; also BGPU_SPH_Z5=0)
  > 121 columns           the general per-particle loops   (kernels.cu; also BGPU_SPH_COLS=0)
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def compare(monkeypatch, h_rel, rsd, env):
    from barcode_b200.chain import Chain, Params
    from oracle import barcode_oracle as bo
    for k in ("BGPU_SPH_Z5", "BGPU_SPH_COLS"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)   # read by bgpu_create
    N, L = 16, 50.0
    kw = dict(N1=N, L1=L, masskernel=3, likelihood=1, rsd_model=rsd, calc_h=2, mass_type=1, sfmodel=1,
              particle_kernel_h_rel=h_rel)
    rng = np.random.default_rng(int(100 * h_rel) + rsd)
    kk = np.fft.fftfreq(N, d=L / N) * 2 * np.pi
    k2 = kk[:, None, None] ** 2 + kk[None, :, None] ** 2 + kk[None, None, :] ** 2
    power = np.where(k2 > 0, 2.0e3 / (1.0 + (np.sqrt(k2) / 0.2) ** 2), 0.0)
    # a smooth field with displacements of about a cell, so that particles leave their cells and wrap around the box
    white = rng.standard_normal((N, N, N))
    s = np.fft.ifftn(np.fft.fftn(white) * np.sqrt(power / power.max())).real
    s *= 0.6 / s.std()
    nobs = np.maximum(0.0, 1.0 + 0.5 * rng.standard_normal((N, N, N)))
    ones = np.ones((N, N, N))
    with Chain(Params(**kw)) as ch:
        ch.set_static(Power=power, nobs=nobs, noise=ones, window=ones)
        pp, pl, dX = ch.psi(s)
        g = ch.gradient_psi(s)
    p = bo.Params(**kw)
    ppo, plo, dXo = bo.psi(p, s, power, nobs, ones, ones)
    go = bo.gradient_psi(p, s, power, nobs, ones, ones)
    assert rel_l2(dX, dXo) < 1e-10
    assert abs(pl - plo) <= 1e-10 * abs(plo) and abs(pp - ppo) <= 1e-10 * abs(ppo)
    assert rel_l2(g, go) < 1e-10


@pytest.mark.parametrize("h_rel", [0.8, 1.0, 1.2, 1.7, 2.6])
@pytest.mark.parametrize("rsd", [False, True])
def test_sph_kernel_and_adjoint_for_every_hull_width(monkeypatch, h_rel, rsd):
    compare(monkeypatch, h_rel, rsd, {})


@pytest.mark.parametrize("env", [{"BGPU_SPH_Z5": "0"}, {"BGPU_SPH_COLS": "0"}])
def test_sph_fallback_paths_at_the_default_width(monkeypatch, env):
    compare(monkeypatch, 1.0, True, env)

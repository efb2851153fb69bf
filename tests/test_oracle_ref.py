"""The compiled reference (oracle/_ref, built from /root/reference by oracle/Makefile) against
the numpy restatement on fresh seeded inputs, plus checks of the FFTW / GSL shims it is linked
against.  Skipped where the library has not been built.  Runs on CPU."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import barcode_oracle as bo
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libbarcode_ref.so not built")


@pytest.mark.parametrize("n", [8, 16, 32])
def test_fft_shim_matches_numpy(n):
    a = np.random.default_rng(n).standard_normal((n, n, n))
    assert rel_l2(ref.fft_r2c(a), np.fft.rfftn(a)) < 1e-14
    assert rel_l2(ref.fft_c2r(np.fft.rfftn(a)), a) < 1e-14


def test_gsl_shim_raw_stream_is_mt19937():
    from numpy.random import MT19937
    for seed in (1, 7, 4357):
        raw, gauss = ref.rng_stream(seed, 64, 64)
        bg = MT19937()
        bg._legacy_seeding(seed)
        assert np.array_equal(raw, bg.random_raw(64))
        assert np.allclose(gauss, bo.gsl_mt19937_gaussians(seed, 64), rtol=4e-16, atol=0)


def test_gradient_operators():
    N, L = 16, 50.0
    a = np.random.default_rng(2).standard_normal((N, N, N))
    p = bo.Params(N1=N, L1=L)
    for dim in (1, 2, 3):
        assert rel_l2(bo.gradfft(p, a, dim), ref.gradfft(a, L, dim)) < 1e-13
        assert rel_l2(bo.gradfindif(p, a, dim), ref.gradfindif(a, L, dim)) < 1e-14


@pytest.mark.parametrize("mk,like,rsd,calc_h,mass_type", [
    (1, 1, False, 0, 1), (1, 1, True, 0, 1), (2, 0, False, 0, 1), (2, 1, True, 1, 0), (0, 1, False, 1, 4),
    (1, 0, True, 0, 1),
])
def test_reference_vs_restatement(mk, like, rsd, calc_h, mass_type):
    from barcode_b200 import inputs
    N, L = 16, 50.0
    cfg = ref.Config(N1=N, L1=L, masskernel=mk, likelihood=like, rsd_model=rsd, calc_h=calc_h, mass_type=mass_type,
                     N_eps_fac=8.0, eps_fac=1.0)
    R = ref.Reference(cfg)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L).ravel()
    rng = np.random.default_rng(100 + mk + 10 * like)
    one = np.ones(R.N)
    R.set_inputs(Power=P, window=one, noise=one, nobs=one)
    truth = R.create_garfield(11, P)
    dX = R.forward(truth, want_pos=False)
    nobs = np.maximum(0, 1 + dX + rng.standard_normal(R.N)) if like == 1 else rng.poisson(np.maximum(1 + dX, 0)) * 1.0
    noise = 1.0 + 0.5 * rng.random(R.N)
    R.set_inputs(nobs=nobs, noise=noise)
    s = 0.5 * R.create_garfield(12, P)
    mf, mr = R.hamiltonian_mass()
    p = bo.Params(N1=N, L1=L, masskernel=mk, likelihood=like, rsd_model=rsd, calc_h=calc_h, mass_type=mass_type,
                  D1=R.scalar("D1"))
    tol = 1e-6 if mk == 0 else 1e-10
    assert abs(bo.fgrow(1.0, p.OM, p.OL) - R.scalar("fgrow")) < 1e-15
    assert rel_l2(bo.gradient_psi(p, s, P, nobs, noise, one), R.gradient_psi(s)) < tol
    pp, pl = R.psi(s)
    ppo, plo, dXo = bo.psi(p, s, P, nobs, noise, one)
    assert abs(pp - ppo) <= 1e-12 * abs(pp) and abs(pl - plo) <= tol * abs(pl)
    # momenta: same mt19937 stream, shell order, colouring
    mom = R.draw_momenta(5)
    W = bo.white_noise_shell_order(N, bo.gsl_mt19937_gaussians(5, 2 * N ** 3)) if p.mass_fs else None
    gauss = None
    if p.mass_rs:
        gauss = bo.gsl_mt19937_gaussians(5, N ** 3)
    momo = bo.draw_momenta(p, W, mf.reshape(N, N, N), mr.reshape(N, N, N), gauss)
    assert rel_l2(momo, mom) < 1e-13
    assert abs(bo.kinetic_term(p, mom, mf, mr) - R.kinetic(mom)) <= 1e-12 * abs(R.kinetic(mom))
    R.close()


def test_reference_smoke_config_runs(tmp_path):
    """The reference's only integration test (test/run/input.par via .travis.yml:75-80: its shipped data/input.par
    at Nx = 8, Lx = 500, N_Gibbs = 5 -- SPH kernel, calc_h = 2, adaptive step size; passes if the process exits).
    Here: the unmodified reference program built by oracle/Makefile runs it to completion and writes the log and
    the per-sample fields; tests/test_dropin.py compares the GPU drop-in against exactly this run."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    td = importlib.import_module("test_dropin")
    if not os.path.exists(td.CPU):
        pytest.skip("oracle/_ref/barcode_cpu not built")
    with np.load(os.path.join(td.ROOT, "barcode_b200", "data", "pk_table.npz")) as f:
        k, P = f["k"], f["P"]
    pk = tmp_path / "pk.dat"
    with open(pk, "w") as o:
        for a, b in zip(k, P):
            o.write(f"{a:.9g} {b:.9g}\n")
    par = td.make_par(calc_h=2, rsd="false", likelihood=1, eps_fac=0.0, mass_type=1, pk=pk, N=8, L=500.0, n_gibbs=5,
                      masskernel=3, n_eps_fac=8.0, eps_update=3, n_bin=200)
    hdr, log = td.run(td.CPU, str(tmp_path / "cpu"), par)
    assert hdr.split("\t")[:3] == ["accepted", "epsilon", "Neps"]
    assert log.shape[1] == 14 and log[:, 0].sum() == 5 and np.all(np.isfinite(log))
    for name in ("deltaLAG_5", "deltaEUL_5", "auxmass_f", "nobs"):
        assert np.fromfile(tmp_path / "cpu" / "data" / name).shape == (8 ** 3,)

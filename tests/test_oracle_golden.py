"""The numpy restatement (oracle/barcode_oracle.py) against the golden vectors that the
compiled, unmodified reference produced (tests/golden/make_golden.py).  Runs on CPU."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2
from oracle import barcode_oracle as bo

TOL = 1e-10


def params(cfg, **over):
    kw = dict(N1=cfg["N1"], L1=cfg["L1"], masskernel=cfg["masskernel"], likelihood=cfg["likelihood"],
              rsd_model=cfg["rsd_model"], calc_h=cfg["calc_h"], mass_type=cfg["mass_type"],
              deltaQ_factor=cfg.get("deltaQ_factor", 1.0), mass_factor=cfg.get("mass_factor", 1.0),
              D1=1.0, sfmodel=cfg.get("sfmodel", 1), slength=cfg.get("slength", 4.0),
              particle_kernel_h_rel=cfg.get("particle_kernel_h_rel", 1.0),
              delta_min=cfg.get("delta_min", -0.999), N_bin=cfg.get("N_bin", 200))
    kw.update(over)
    return bo.Params(**kw)


def test_power_table_matches_reference_readtab(case):
    from barcode_b200 import inputs
    k_tab, p_tab = inputs.load_pk_table()
    P = bo.power_on_grid(k_tab, p_tab, case["cfg"]["N1"], case["cfg"]["L1"])
    assert np.array_equal(P.ravel(), case["Power"])          # bit for bit
    assert np.array_equal(inputs.power_on_grid(k_tab, p_tab, case["cfg"]["N1"], case["cfg"]["L1"]).ravel(),
                          case["Power"])


def test_forward_positions_and_density(case):
    p = params(case["cfg"])
    N = p.N1
    dX, (x, y, z), _ = bo.forward(p, case["signal"].reshape(N, N, N))
    for a, b in ((x, case["posx"]), (y, case["posy"]), (z, case["posz"])):
        dd = np.abs(a.ravel() - b)
        assert np.minimum(dd, p.L1 - dd).max() < 1e-11
    if p.masskernel == 0:
        assert np.count_nonzero(np.abs(dX.ravel() - case["deltaX_fwd"]) > 1e-9) <= 4
    else:
        assert rel_l2(dX, case["deltaX_fwd"]) < TOL


def test_density_on_reference_positions(case):
    p = params(case["cfg"])
    rho = bo.density(p, case["posx"], case["posy"], case["posz"])
    assert rel_l2(bo.overdens(rho), case["deltaX_fwd"]) < 1e-13


def test_gradients(case):
    p = params(case["cfg"])
    args = (case["signal"], case["nobs"], case["noise"], case["window"])
    tol = 1e-6 if p.masskernel == 0 else TOL
    assert rel_l2(bo.grad_log_prior(p, case["signal"], case["Power"]), case["grad_prior"]) < TOL
    # (the harness's grad_like entry calls likelihood_grad_log_like, which gradient_psi bypasses for the
    # Gaussian-random-field likelihood, HMC.cc:159-160: there the reference's split is gradpsi - grad_prior)
    like_ref = case["grad_like"] if p.likelihood != 3 else case["gradpsi"] - case["grad_prior"]
    assert rel_l2(bo.grad_log_like(p, *args)[0], like_ref) < tol
    assert rel_l2(bo.gradient_psi(p, case["signal"], case["Power"], *args[1:]), case["gradpsi"]) < tol


def test_energies(case):
    p = params(case["cfg"])
    pp, pl, dX = bo.psi(p, case["signal"], case["Power"], case["nobs"], case["noise"], case["window"])
    assert abs(pp - case["psi_prior"]) <= TOL * abs(case["psi_prior"])
    tol = 1e-6 if p.masskernel == 0 else TOL
    assert abs(pl - case["psi_like"]) <= tol * abs(case["psi_like"])
    mf, mr = bo.hamiltonian_mass(p, case["Power"], case["signal"], case["nobs"], case["noise"], case["window"])
    if p.mass_type in (2, 3):   # a binned spectrum is in the mass: equal to summation order, not bit for bit
        assert rel_l2(mf, case["mass_f"]) < 1e-13
    else:
        assert np.array_equal(mf.ravel(), case["mass_f"])
    assert np.array_equal(mr.ravel(), case["mass_r"])
    K = bo.kinetic_term(p, case["momenta"], mf, mr)
    assert abs(K - case["K"]) <= TOL * abs(case["K"])


def test_leapfrog_and_delta_H(case):
    p = params(case["cfg"])
    if p.masskernel == 0:
        pytest.skip("NGP density is discontinuous in the displacement")
    mf, mr = case["mass_f"], case["mass_r"]
    args = (case["Power"], case["nobs"], case["noise"], case["window"], mf, mr)
    sf, pf = bo.leapfrog(p, case["signal"], case["momenta"], int(case["Neps"]), float(case["epsilon"]), *args)
    assert rel_l2(sf, case["s_f"]) < 1e-8 and rel_l2(pf, case["p_f"]) < 1e-8
    dH, sc, _ = bo.delta_hamiltonian(p, case["signal"], case["momenta"], case["s_f"], case["p_f"], *args)
    assert abs(dH - case["dH"]) <= 1e-8 * abs(case["dH"])
    for k in ("H_kin_i", "H_kin_f", "psi_prior_i", "psi_prior_f", "psi_likeli_i", "psi_likeli_f"):
        assert abs(sc[k] - case["dh_" + k]) <= 1e-9 * abs(case["dh_" + k])


def test_momentum_draw_golden():
    with np.load(os.path.join(GOLDEN, "garfield_n8.npz")) as f:
        g = {k: f[k] for k in f.files}
    N = 8
    # GSL mt19937 + polar Gaussian stream
    gauss = bo.gsl_mt19937_gaussians(7, 2 * N ** 3)
    # numpy's vectorised log/sqrt may differ from libm in the last bit
    assert np.allclose(gauss[:16], g["gauss"], rtol=4e-16, atol=0)
    W = bo.white_noise_shell_order(N, gauss)
    assert np.allclose(W, g["white"], rtol=4e-16, atol=0)
    p = bo.Params(N1=N, L1=25.0, mass_type=1)
    assert rel_l2(bo.create_garfield(p, W, g["Power"]), g["field"]) < 1e-13
    assert rel_l2(bo.draw_momenta(p, W, g["mass_f"].reshape(N, N, N), None), g["momenta"]) < 1e-13


def test_mt19937_raw_stream_golden():
    from numpy.random import MT19937
    with np.load(os.path.join(GOLDEN, "garfield_n8.npz")) as f:
        raw = f["raw"]
    bg = MT19937()
    bg._legacy_seeding(7)
    assert np.array_equal(bg.random_raw(16), raw)


def test_exact_adjoint_is_the_derivative_of_psi():
    """calc_h = 4 (new): central finite difference of the oracle's own psi()."""
    from conftest import load_case
    for name in ("za_cic_gauss", "za_tsc_gauss_rsd_mass0"):
        c = load_case(name)
        p = params(c["cfg"], calc_h=4)
        N = p.N1
        one = np.ones_like(c["window"])
        s = c["signal"].reshape(N, N, N)
        g = bo.gradient_psi(p, s, c["Power"], c["nobs"], c["noise"], one)
        v = np.random.default_rng(3).standard_normal(s.shape)
        v *= 1e-6 / np.abs(v).max()
        f = lambda q: sum(bo.psi(p, q, c["Power"], c["nobs"], c["noise"], one)[:2])
        fd = (f(s + v) - f(s - v)) / 2.0
        an = float(np.sum(g * v))
        assert abs(fd - an) <= 2e-5 * abs(an) + 1e-9


def test_measure_spectrum_golden():
    """measure_spectrum (field_statistics.cpp:20-90): the numpy restatement against the compiled reference."""
    with np.load(os.path.join(GOLDEN, "spectrum_n16.npz")) as f:
        g = {k: f[k] for k in f.files}
    p = bo.Params(N1=16, L1=float(g["L1"]))
    for nb in (20, 200):
        km, pw = bo.measure_spectrum(p, g["signal"], nb)
        assert np.allclose(km, g[f"kmode_{nb}"], rtol=1e-13, atol=0)
        assert np.allclose(pw, g[f"power_{nb}"], rtol=1e-12, atol=0)

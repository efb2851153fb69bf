"""Parity of the CUDA path (through the C ABI) with the reference.

Checked against (a) the golden vectors the compiled reference produced
(tests/golden/, generator make_golden.py) and (b) the numpy restatement
(oracle/barcode_oracle.py) on seeded inputs.  Tolerances are BASELINE.json's:
integer cell indices bit-exact; density, log-posterior and gradient within
1e-10 relative L2 (FP64); Delta-H over the fixed trajectory within 1e-8 relative.
"""
import os

import numpy as np
import pytest

from conftest import CASE_NAMES, load_case, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-10


def make_chain(cfg, **over):
    from barcode_b200.chain import Chain, Params
    kw = dict(N1=cfg["N1"], L1=cfg["L1"], masskernel=cfg["masskernel"], likelihood=cfg["likelihood"],
              rsd_model=cfg["rsd_model"], calc_h=cfg["calc_h"], mass_type=cfg["mass_type"],
              sfmodel=cfg.get("sfmodel", 1), deltaQ_factor=cfg.get("deltaQ_factor", 1.0),
              mass_factor=cfg.get("mass_factor", 1.0), slength=cfg.get("slength", 4.0),
              particle_kernel_h_rel=cfg.get("particle_kernel_h_rel", 1.0),
              delta_min=cfg.get("delta_min", -0.999), N_bin=cfg.get("N_bin", 200))
    kw.update(over)
    return Chain(Params(**kw))


def oracle_params(cfg, **over):
    from oracle import barcode_oracle as bo
    kw = dict(N1=cfg["N1"], L1=cfg["L1"], masskernel=cfg["masskernel"], likelihood=cfg["likelihood"],
              rsd_model=cfg["rsd_model"], calc_h=cfg["calc_h"], mass_type=cfg["mass_type"],
              deltaQ_factor=cfg.get("deltaQ_factor", 1.0), mass_factor=cfg.get("mass_factor", 1.0),
              sfmodel=cfg.get("sfmodel", 1), slength=cfg.get("slength", 4.0),
              particle_kernel_h_rel=cfg.get("particle_kernel_h_rel", 1.0),
              delta_min=cfg.get("delta_min", -0.999), N_bin=cfg.get("N_bin", 200))
    kw.update(over)
    return bo.Params(**kw)


def loaded_chain(c, **over):
    ch = make_chain(c["cfg"], **over)
    ch.set_static(Power=c["Power"], nobs=c["nobs"], noise=c["noise"], window=c["window"])
    ch.set_mass(mass_f=c["mass_f"], mass_r=c["mass_r"])
    return ch


# ---------------------------------------------------------------- FFT
@pytest.mark.parametrize("N", [8, 16, 32, 64, 128, 256, 512])
def test_fft_matches_numpy(N):
    from barcode_b200.chain import Chain, Params
    rng = np.random.default_rng(N)
    a = rng.standard_normal((N, N, N))
    with Chain(Params(N1=N, L1=100.0)) as ch:
        c = ch.fft_r2c(a)
        ref = np.fft.rfftn(a)
        assert rel_l2(c, ref) < 1e-14
        back = ch.fft_c2r(ref)
        assert rel_l2(back, a) < 1e-14


def test_convolve_inv_corr(case):
    from oracle import barcode_oracle as bo
    with loaded_chain(case) as ch:
        out = ch.convolve_inv_corr(case["signal"], case["Power"])
    assert rel_l2(out, case["grad_prior"]) < TOL
    p = oracle_params(case["cfg"])
    assert rel_l2(out, bo.convolve_inv_corr(p, case["signal"], case["Power"])) < TOL


# ---------------------------------------------------------------- particles
def test_cell_indices_bit_exact(case):
    from oracle import barcode_oracle as bo
    cfg = case["cfg"]
    p = oracle_params(cfg)
    x, y, z = case["posx"], case["posy"], case["posz"]
    with make_chain(cfg) as ch:
        ci, cj, ck = ch.cell_indices(x, y, z)
    if cfg["masskernel"] == 1:
        ri, rj, rk = (bo.cic_cells_weights(p, a)[0] for a in (x, y, z))
    else:
        ri, rj, rk = bo.ngp_cells(p, x, p.min1), bo.ngp_cells(p, y, p.min2), bo.ngp_cells(p, z, p.min3)
    assert np.array_equal(ci, ri) and np.array_equal(cj, rj) and np.array_equal(ck, rk)


def test_cell_indices_edges():
    """positions on cell faces, at 0, just below L, half-cell offsets"""
    from barcode_b200.chain import Chain, Params
    from oracle import barcode_oracle as bo
    N, L = 16, 50.0
    d = L / N
    vals = np.array([0.0, np.nextafter(L, 0), 0.5 * d, np.nextafter(0.5 * d, 0), np.nextafter(0.5 * d, 1), d,
                     np.nextafter(d, 0), 7 * d, 7.5 * d, L - 0.5 * d, np.nextafter(L - 0.5 * d, 0), 15.9999999 * d])
    x, y, z = [a.ravel() for a in np.meshgrid(vals, vals, vals, indexing="ij")]
    for mk in (0, 1, 2):
        p = bo.Params(N1=N, L1=L, masskernel=mk)
        with Chain(Params(N1=N, L1=L, masskernel=mk)) as ch:
            ci, cj, ck = ch.cell_indices(x, y, z)
        if mk == 1:
            r = [bo.cic_cells_weights(p, a)[0] for a in (x, y, z)]
        else:
            r = [bo.ngp_cells(p, a, 0.0) for a in (x, y, z)]
        assert np.array_equal(ci, r[0]) and np.array_equal(cj, r[1]) and np.array_equal(ck, r[2])


def test_density_same_positions(case):
    from oracle import barcode_oracle as bo
    cfg = case["cfg"]
    with make_chain(cfg) as ch:
        rho = ch.assign_density(case["posx"], case["posy"], case["posz"])
    ref = bo.density(oracle_params(cfg), case["posx"], case["posy"], case["posz"])
    assert rel_l2(rho, ref) < 1e-13
    if cfg["masskernel"] != 3:  # the SPH kernel is not normalised on the grid (massFunctions.cc:476-494)
        assert abs(rho.sum() - rho.size) < 1e-8 * rho.size  # unit masses, partition of unity


def test_forward_density_and_positions(case):
    with loaded_chain(case) as ch:
        dX, x, y, z = ch.forward(case["signal"], want_pos=True)
    L = case["cfg"]["L1"]
    for a, b in ((x, case["posx"]), (y, case["posy"]), (z, case["posz"])):
        dd = np.abs(a - b)
        dd = np.minimum(dd, L - dd)  # a particle within rounding of the box edge may wrap
        assert dd.max() < 1e-11
    if case["cfg"]["masskernel"] == 0:
        # NGP is discontinuous: allow the handful of particles that sit within rounding of a cell face
        assert np.count_nonzero(np.abs(dX.ravel() - case["deltaX_fwd"]) > 1e-9) <= 4
    else:
        assert rel_l2(dX, case["deltaX_fwd"]) < TOL


# ---------------------------------------------------------------- posterior
def test_gradient_psi(case):
    with loaded_chain(case) as ch:
        g = ch.gradient_psi(case["signal"])
    if case["cfg"]["masskernel"] == 0:
        assert rel_l2(g, case["gradpsi"]) < 1e-6
    else:
        assert rel_l2(g, case["gradpsi"]) < TOL


def test_psi_and_deltaX(case):
    with loaded_chain(case) as ch:
        pp, pl, dX = ch.psi(case["signal"])
    assert abs(pp - case["psi_prior"]) <= TOL * abs(case["psi_prior"])
    if case["cfg"]["masskernel"] == 0:
        assert abs(pl - case["psi_like"]) <= 1e-6 * abs(case["psi_like"])
    else:
        assert abs(pl - case["psi_like"]) <= TOL * abs(case["psi_like"])
        if case["cfg"]["likelihood"] != 3:  # the GRF likelihood runs no forward model: deltaX is not refreshed
            assert rel_l2(dX, case["deltaX_psi"]) < TOL


def test_kinetic_term(case):
    with loaded_chain(case) as ch:
        K = ch.kinetic_term(case["momenta"])
    assert abs(K - case["K"]) <= TOL * abs(case["K"])


def test_hamiltonian_mass(case):
    with make_chain(case["cfg"]) as ch:
        if case["cfg"]["mass_type"] in (2, 3):
            # likelihood-force masses (HMC_mass.cc:39-160): gradient of the likelihood at the signal, its binned
            # spectrum, 2/P + sqrt(F/P) -- equal to the order of the spectrum's sums
            ch.set_static(Power=case["Power"], nobs=case["nobs"], noise=case["noise"], window=case["window"])
            mf, mr = ch.hamiltonian_mass(case["signal"])
            assert rel_l2(mf, case["mass_f"]) < TOL
            return
        ch.set_static(Power=case["Power"])
        mf, mr = ch.hamiltonian_mass()
    assert np.array_equal(mf.ravel(), case["mass_f"]) and np.array_equal(mr.ravel(), case["mass_r"])


def test_leapfrog_and_delta_H(case):
    if case["cfg"]["masskernel"] == 0:
        pytest.skip("NGP density is discontinuous in the displacement: trajectories are not comparable at 1e-8")
    with loaded_chain(case) as ch:
        sf, pf = ch.leapfrog(case["signal"], case["momenta"], int(case["Neps"]), float(case["epsilon"]))
        assert rel_l2(sf, case["s_f"]) < 1e-8
        assert rel_l2(pf, case["p_f"]) < 1e-8
        dH, sc, _ = ch.delta_hamiltonian(case["signal"], case["momenta"], sf, pf)
    # 1e-8 relative (BASELINE.json) plus the FP64 rounding floor of the terms dH is the difference of
    # (matters only for the gentle log-normal trajectories, where dH ~ 1e-2 while K ~ 1e9)
    floor = 1e-13 * sum(abs(float(case["dh_" + k])) for k in ("H_kin_i", "psi_prior_i", "psi_likeli_i"))
    assert abs(dH - case["dH"]) <= 1e-8 * abs(case["dH"]) + floor
    for k in ("H_kin_i", "H_kin_f", "psi_prior_i", "psi_prior_f", "psi_likeli_i", "psi_likeli_f"):
        assert abs(sc[k] - case["dh_" + k]) <= 1e-8 * abs(case["dh_" + k]), k


# ---------------------------------------------------------------- momenta
def test_colour_momenta_golden():
    import os
    from conftest import GOLDEN
    from barcode_b200.chain import Chain, Params
    with np.load(os.path.join(GOLDEN, "garfield_n8.npz")) as f:
        g = {k: f[k] for k in f.files}
    with Chain(Params(N1=8, L1=25.0, mass_type=4)) as ch:
        ch.set_mass(mass_f=g["Power"])
        field = ch.color_momenta(g["white"])
        assert rel_l2(field, g["field"]) < 1e-13
    with Chain(Params(N1=8, L1=25.0, mass_type=1)) as ch:
        ch.set_mass(mass_f=g["mass_f"])
        mom = ch.color_momenta(g["white"])
        assert rel_l2(mom, g["momenta"]) < 1e-13


@pytest.mark.parametrize("N", [16, 32])
def test_colour_momenta_oracle(N):
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    from oracle import barcode_oracle as bo
    L = inputs.box_length(N)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    W = inputs.complex_white_noise(N, 5)
    p = bo.Params(N1=N, L1=L, mass_type=4)
    with Chain(Params(N1=N, L1=L, mass_type=4)) as ch:
        ch.set_mass(mass_f=P)
        out = ch.color_momenta(W)
    assert rel_l2(out, bo.create_garfield(p, W, P)) < 1e-13


# ---------------------------------------------------------------- exact adjoint (new): finite differences
@pytest.mark.parametrize("name", ["za_cic_gauss", "za_cic_gauss_rsd", "za_tsc_poisson", "za_tsc_gauss_rsd_mass0",
                                  "alpt_cic_gauss", "alpt_tsc_poisson_h1"])
def test_exact_adjoint_matches_oracle_and_fd(name):
    """calc_h = 4: the reference has no CIC/TSC adjoint, so the check is the author's own
    (HMC_models.cc:426-431): a central finite difference of psi() along a random direction."""
    from oracle import barcode_oracle as bo
    c = load_case(name)
    cfg = dict(c["cfg"], calc_h=4)
    one = np.ones_like(c["window"])  # binary/unit window: the residual convention drops a factor w
    N = cfg["N1"]
    with make_chain(cfg) as ch:
        ch.set_static(Power=c["Power"], nobs=c["nobs"], noise=c["noise"], window=one)
        s = c["signal"].reshape(N, N, N)
        g = ch.gradient_psi(s)
        p = oracle_params(cfg)
        go = bo.gradient_psi(p, s, c["Power"], c["nobs"], c["noise"], one)
        assert rel_l2(g, go) < TOL
        # the reference's Poisson value ignores RSD and deltaQ_factor while its gradient applies them
        # (poissonian.cpp:54-56): psi is then not the function the gradient belongs to
        if cfg["likelihood"] == 0 and (cfg["rsd_model"] or cfg.get("deltaQ_factor", 1.0) != 1.0):
            return
        rng = np.random.default_rng(3)
        v = rng.standard_normal(s.shape)
        v *= 1e-6 / np.abs(v).max()
        if cfg["likelihood"] == 0:
            psi = lambda q: sum(ch.psi(q, want_deltaX=False)[:2])
        else:
            psi = lambda q: sum(ch.psi(q, want_deltaX=False)[:2])
        fd = (psi(s + v) - psi(s - v)) / 2.0
        an = float(np.sum(g * v))
        assert abs(fd - an) <= 2e-5 * abs(an) + 1e-9


# ---------------------------------------------------------------- size-independent properties at bench sizes
@pytest.mark.parametrize("N", [256])
def test_large_grid_properties(N):
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    L = inputs.box_length(N)
    rng = np.random.default_rng(1)
    with Chain(Params(N1=N, L1=L, masskernel=1, rsd_model=True, calc_h=0)) as ch:
        a = rng.standard_normal((N, N, N))
        # FFT round trip and Parseval
        c = ch.fft_r2c(a)
        back = ch.fft_c2r(c)
        assert rel_l2(back, a) < 1e-14
        w = np.full(c.shape[-1], 2.0)
        w[0] = w[-1] = 1.0
        assert abs(np.sum(w * np.abs(c) ** 2) / a.size - np.sum(a * a)) < 1e-10 * np.sum(a * a)
        prob = inputs.synthetic_problem(ch, seed=1)
        s = prob["signal"]
        dX = ch.forward(s)
        # mass conservation: mean overdensity is zero, rho >= 0
        assert abs(dX.mean()) < 1e-12 and dX.min() >= -1.0 - 1e-12
        # prior gradient is linear
        P = prob["Power"]
        g1 = ch.convolve_inv_corr(s, P)
        g2 = ch.convolve_inv_corr(2.5 * s, P)
        assert rel_l2(g2, 2.5 * g1) < 1e-13
        # leapfrog: energy error shrinks ~ eps^2 and the trajectory is reversible
        p0 = prob["momenta"]
        errs = []
        for eps in (2e-3, 1e-3):
            sf, pf = ch.leapfrog(s, p0, 2, eps)
            dH, _, _ = ch.delta_hamiltonian(s, p0, sf, pf)
            errs.append(abs(dH))
        assert errs[1] < errs[0]
        sb, pb = ch.leapfrog(sf, -pf, 2, 1e-3)
        assert rel_l2(sb, s) < 1e-9 and rel_l2(pb, -p0) < 1e-9


# ---------------------------------------------------------------- the TMA-staged FFT pass (N >= 128)
def _problem_128(seed=11):
    from barcode_b200 import inputs
    N = 128
    L = inputs.box_length(N)
    rng = np.random.default_rng(seed)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    nobs = np.maximum(0.0, 1.0 + 0.3 * rng.standard_normal(n))
    noise = np.ones(n)
    window = np.ones(n)
    s = 0.4 * rng.standard_normal((N, N, N))
    return N, L, P, nobs, noise, window, s


@pytest.mark.parametrize("calc_h,masskernel,rsd,sfmodel", [(0, 1, True, 1), (4, 1, False, 1), (0, 2, False, 1),
                                                            (0, 1, False, 2), (4, 1, False, 2)])
def test_gradient_128_matches_oracle(calc_h, masskernel, rsd, sfmodel):
    """128^3 runs through the TMA-staged strided pass (fft_tma.cuh); the oracle is the numpy restatement.
    sfmodel = 2 is Lag2Eul_non_zeldovich (2LPT + spherical collapse, ALPT split at slength = 8 Mpc/h); with
    calc_h = 4 its exact adjoint (new; the oracle's is validated by finite differences of psi)."""
    from barcode_b200.chain import Chain, Params
    from oracle import barcode_oracle as bo
    N, L, P, nobs, noise, window, s = _problem_128()
    kw = dict(N1=N, L1=L, masskernel=masskernel, likelihood=1, rsd_model=rsd, calc_h=calc_h, mass_type=1,
              sfmodel=sfmodel, slength=8.0)
    with Chain(Params(**kw)) as ch:
        ch.set_static(Power=P, nobs=nobs, noise=noise, window=window)
        g = ch.gradient_psi(s)
        pp, pl, dX = ch.psi(s)
    p = bo.Params(**kw)
    go = bo.gradient_psi(p, s, P, nobs, noise, window)
    assert rel_l2(g, go) < TOL
    po, lo, dXo = bo.psi(p, s, P, nobs, noise, window)
    assert abs(pp - po) <= TOL * abs(po) and abs(pl - lo) <= TOL * abs(lo)
    assert rel_l2(dX, dXo) < TOL


@pytest.mark.parametrize("rsd,amp", [(True, 1.0), (False, 1.0), (True, 40.0), (False, 4000.0)])
def test_particle_kernel_generations_agree(rsd, amp, monkeypatch):
    """The three generations of the CIC scatter / gather -- particle per thread (kernels.cu), x sweep, lean x sweep
    (particles_sweep.cu) -- deposit the same products into the same cells: density and exact-adjoint gradient agree to
    summation order.  amp = 40 wraps particles around the box; amp = 4000 throws them several box lengths away, which
    is the lean kernels' out-of-line general path (coordinates beyond 3/4 of a box length from the box)."""
    from barcode_b200.chain import Chain, Params
    N, L, P, nobs, noise, window, s = _problem_128()
    s = amp * s
    out = {}
    for tag, env in (("lean", {}), ("sweep", {"BGPU_LEAN": "0"}), ("thread", {"BGPU_SWEEP": "0"})):
        for k in ("BGPU_LEAN", "BGPU_SWEEP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Chain(Params(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=rsd, calc_h=4, mass_type=1)) as ch:
            ch.set_static(Power=P, nobs=nobs, noise=noise, window=window)
            out[tag] = (ch.forward(s), ch.gradient_psi(s))
    for tag in ("sweep", "thread"):
        assert rel_l2(out["lean"][0], out[tag][0]) < 1e-13, tag
        assert rel_l2(out["lean"][1], out[tag][1]) < 1e-12, tag
    assert abs(out["lean"][0].sum()) < 1e-6 * N ** 3   # delta_x sums to zero: every particle was deposited


@pytest.mark.parametrize("N", [128, 256])
def test_tma_pass_matches_cp_async_pass(N, monkeypatch):
    """Two independent implementations of the strided pass (TMA ring vs cp.async) agree to rounding."""
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    L = inputs.box_length(N)
    rng = np.random.default_rng(5)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    s = 0.3 * rng.standard_normal(n)
    nobs = 1.0 + 0.1 * rng.standard_normal(n)
    out = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("BGPU_FFT_TMA", tma)
        with Chain(Params(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=0)) as ch:
            ch.set_static(Power=P, nobs=nobs, noise=np.ones(n), window=np.ones(n))
            ch.hamiltonian_mass()
            out[tma] = (ch.gradient_psi(s), ch.convolve_inv_corr(s, P), ch.kinetic_term(s))
    assert rel_l2(out["1"][0], out["0"][0]) < 1e-13
    assert rel_l2(out["1"][1], out["0"][1]) < 1e-14
    assert abs(out["1"][2] - out["0"][2]) <= 1e-13 * abs(out["0"][2])


@pytest.mark.skipif(not os.environ.get("BGPU_HEAVY_TESTS"), reason="512^3: two 22 GB chains; set BGPU_HEAVY_TESTS=1 "
                    "(the same comparison runs without Python in tools/native/fft_ab.cc, "
                    "profiles/fft_ab_r01f_2warp_512.log)")
def test_two_warp_pencils_match_one_warp_pencils_512(monkeypatch):
    """BGPU_FFT_2WARP=1 spreads a 512-point strided pencil over two warps (named barrier per pencil,
    fft_tma.cuh ColAccessWide).  Same radix sequence and twiddles, so the transforms are bit-identical."""
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    N = 512
    L = inputs.box_length(N)
    a = np.random.default_rng(9).standard_normal((N, N, N))
    out = {}
    for v in ("0", "1"):
        monkeypatch.setenv("BGPU_FFT_2WARP", v)
        with Chain(Params(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=0)) as ch:
            c = ch.fft_r2c(a)
            out[v] = (c, ch.fft_c2r(c))
    assert np.array_equal(out["0"][0], out["1"][0])
    assert np.array_equal(out["0"][1], out["1"][1])
    assert rel_l2(out["1"][1], a) < 1e-14


@pytest.mark.parametrize("calc_h,sfmodel,rsd,like", [(0, 1, True, 1), (4, 1, True, 1), (4, 1, False, 1), (0, 3, False, 1),
                                                     (4, 3, False, 1), (0, 1, False, 0), (0, 1, False, 2)])
def test_shared_x_pass_matches_separate_transforms(calc_h, sfmodel, rsd, like, monkeypatch):
    """BGPU_SHARE_X=1: the y and z components of the displacement, gradient and back-projection triples share one
    x pass (fft_ops.h K_MULK*, K_COMP_UNIT) -- 9 x passes per calc_h = 0 evaluation instead of 12, same result to rounding."""
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    N = 128
    L = inputs.box_length(N)
    rng = np.random.default_rng(21)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    s = 0.3 * rng.standard_normal(n)
    nobs = 1.0 + 0.1 * rng.standard_normal(n)
    out = {}
    for v in ("0", "1"):
        monkeypatch.setenv("BGPU_SHARE_X", v)
        with Chain(Params(N1=N, L1=L, masskernel=1, likelihood=like, rsd_model=rsd, sfmodel=sfmodel, calc_h=calc_h,
                          correct_delta=True)) as ch:
            ch.set_static(Power=P, nobs=nobs, noise=np.ones(n), window=np.ones(n))
            out[v] = (ch.gradient_psi(s), ch.forward(s), ch.psi(s)[:2])
    assert rel_l2(out["1"][1], out["0"][1]) < 1e-13
    assert rel_l2(out["1"][0], out["0"][0]) < 1e-12
    for a, b in zip(out["1"][2], out["0"][2]):
        assert abs(a - b) <= 1e-12 * abs(b)


def test_fused_zy_kernel_matches_separate_passes(monkeypatch):
    """The opt-in fused z+y kernel (fft_fused.cuh, BGPU_FFT_FUSED=1: warp-specialised roles, the intermediate
    array handed over through L2 with per-plane flags) against numpy and against the separate passes."""
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    N = 256
    L = inputs.box_length(N)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((N, N, N))
    ref = np.fft.rfftn(x)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    s = 0.3 * rng.standard_normal(n)
    nobs = 1.0 + 0.1 * rng.standard_normal(n)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("BGPU_FFT_FUSED", fused)
        with Chain(Params(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=0)) as ch:
            if fused == "1":
                assert rel_l2(ch.fft_r2c(x), ref) < 1e-14
                assert rel_l2(ch.fft_c2r(ref), x) < 1e-14
            ch.set_static(Power=P, nobs=nobs, noise=np.ones(n), window=np.ones(n))
            ch.hamiltonian_mass()
            out[fused] = (ch.gradient_psi(s), ch.kinetic_term(s))
    assert rel_l2(out["1"][0], out["0"][0]) < 1e-13
    assert abs(out["1"][1] - out["0"][1]) <= 1e-13 * abs(out["0"][1])


# ---------------------------------------------------------------- device momentum draw
def test_device_normals_match_oracle_generator():
    """The Philox4x32-10 + Box-Muller stream on the GPU against its numpy restatement (itself pinned to the
    published known-answer vectors, tests/test_host.py): same uniforms bit for bit, normals to libm rounding."""
    from oracle import barcode_oracle as bo
    from barcode_b200.chain import Chain, Params
    with Chain(Params(N1=32, L1=100.0)) as ch:
        for seed, draw, stream, first, n in ((1, 0, 0, 0, 4096), (0xDEADBEEFCAFE, (1 << 40) + 3, 1, 2048, 1024)):
            got = ch.device_normals(seed, draw, stream, first, n)
            want = bo.device_normals(seed, draw, stream, first, n)
            assert np.max(np.abs(got - want)) < 1e-13


@pytest.mark.parametrize("mass_type", [1, 4, 0])
def test_device_momentum_draw_has_the_mass_as_covariance(mass_type):
    """bgpu_draw_momenta_device: E[K] = (modes with mass)/2 with K from the parity-tested kinetic term, the draws
    are reproducible per (seed, index) and differ between indices, the DC mode is zero (random.cpp:347-351)."""
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    N = 64
    L = inputs.box_length(N)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    with Chain(Params(N1=N, L1=L, mass_type=mass_type)) as ch:
        ch.set_static(Power=P)
        ch.hamiltonian_mass()
        p0 = ch.draw_momenta_device(5, 0)
        assert np.array_equal(p0, ch.draw_momenta_device(5, 0))
        p1 = ch.draw_momenta_device(5, 1)
        assert not np.array_equal(p0, p1)
        nmodes = N ** 3 - (0 if mass_type == 0 else 1)
        Ks = [ch.kinetic_term(ch.draw_momenta_device(5, i)) for i in range(4)]
        # K is chi^2_nmodes / 2: mean nmodes/2, sigma sqrt(nmodes/2); four draws -> 5 sigma of the mean
        assert abs(np.mean(Ks) - nmodes / 2) < 5 * np.sqrt(nmodes / 2) / 2
        if mass_type != 0:
            assert abs(p0.sum()) < 1e-6 * np.abs(p0).sum()
            # binned spectrum follows the mass: <|p^|^2> = N^2/V M
            ph = np.abs(np.fft.rfftn(p0.reshape(N, N, N))) ** 2
            M = ch.hamiltonian_mass()[0].reshape(N, N, N)[:, :, :N // 2 + 1]
            ok = M > 0
            ratio = ph[ok] / (float(N) ** 6 / L ** 3 * M[ok])
            assert abs(ratio.mean() - 1) < 0.02


# ---------------------------------------------------------------- P(k) diagnostic
def test_measure_spectrum_matches_reference_and_oracle():
    import os
    from conftest import GOLDEN
    from oracle import barcode_oracle as bo
    from barcode_b200.chain import Chain, Params
    with np.load(os.path.join(GOLDEN, "spectrum_n16.npz")) as f:
        g = {k: f[k] for k in f.files}
    with Chain(Params(N1=16, L1=float(g["L1"]))) as ch:
        for nb in (20, 200):
            km, pw = ch.measure_spectrum(g["signal"], nb)
            assert np.allclose(km, g[f"kmode_{nb}"], rtol=1e-12, atol=0)
            assert np.allclose(pw, g[f"power_{nb}"], rtol=1e-11, atol=0)
    N, L = 128, 400.0
    x = np.random.default_rng(8).standard_normal((N, N, N))
    with Chain(Params(N1=N, L1=L)) as ch:
        km, pw = ch.measure_spectrum(x, 200)
    km0, pw0 = bo.measure_spectrum(bo.Params(N1=N, L1=L), x, 200)
    assert np.allclose(km, km0, rtol=1e-12, atol=0) and np.allclose(pw, pw0, rtol=1e-11, atol=0)
    # white noise of unit variance: P = V / N^3 in every bin
    assert abs(pw[20:150].mean() / (L ** 3 / N ** 3) - 1) < 0.01


# ---------------------------------------------------------------- BASELINE.json configs[0] against the LIVE reference
@pytest.mark.parametrize("N,mk,like,rsd,calc_h,mass_type,sfmodel", [
    (64, 1, 1, False, 0, 1, 1),     # configs[0]: 64^3 ZA + CIC, Gaussian, real space
    (64, 1, 1, True, 0, 1, 2),      # configs[1]'s physics at 64^3 (RSD => the reference runs Zel'dovich)
    (64, 2, 0, False, 0, 1, 1),     # configs[2]'s physics at 64^3 (TSC, Poisson)
    (64, 3, 1, False, 2, 1, 1),     # the shipped default: SPH + its exact adjoint
    (128, 2, 0, False, 0, 1, 1),    # configs[2] at its own size: 128^3 ZA + TSC, Poisson
    (256, 1, 1, True, 0, 1, 2),     # configs[1] at its own size: 256^3 (2LPT requested ->) ZA + CIC, Gaussian, RSD
    (128, 1, 1, False, 0, 1, 3),    # configs[3]'s physics at 128^3: ALPT + CIC (Lag2Eul_non_zeldovich)
    (512, 1, 1, True, 0, 1, 2),     # the north star's own size: 512^3 ZA + CIC, Gaussian, RSD (no trajectory: CPU minutes)
    pytest.param(512, 1, 1, False, 0, 1, 3, marks=pytest.mark.skipif(
        not os.environ.get("BGPU_HEAVY_TESTS"), reason="configs[3] at its own size, 512^3 ALPT + CIC: minutes of "
        "reference CPU time; set BGPU_HEAVY_TESTS=1 (run once per round, log under profiles/)")),
])
def test_against_the_compiled_reference_at_baseline_sizes(N, mk, like, rsd, calc_h, mass_type, sfmodel, tmp_path,
                                                          monkeypatch):
    """The CUDA path against oracle/_ref (the unmodified reference sources compiled in the build container; the .so
    travels with the snapshot) at the sizes of BASELINE.json configs[0-2] (64^3 is also the reference's own
    data/input.par): forward density, both energies, the gradient, the kinetic energy, a 3-step trajectory and dH."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libbarcode_ref.so not built")
    from barcode_b200 import inputs
    from barcode_b200.chain import Chain, Params
    L = inputs.box_length(N)
    cfg = ref.Config(N1=N, L1=L, masskernel=mk, likelihood=like, rsd_model=rsd, calc_h=calc_h, mass_type=mass_type,
                     sfmodel=sfmodel, N_eps_fac=8.0, eps_fac=1.0)
    R = ref.Reference(cfg)
    if sfmodel != 1 and not rsd:
        # the reference re-reads its ALPT kernel file from the working directory on every forward evaluation
        monkeypatch.chdir(tmp_path)
        R.kernelcomp()
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L).ravel()
    rng = np.random.default_rng(64 + mk)
    one = np.ones(R.N)
    R.set_inputs(Power=P, window=one, noise=one, nobs=one)
    truth = R.create_garfield(21, P)
    dX_truth = R.forward(truth, want_pos=False)
    nobs = np.maximum(0, 1 + dX_truth + rng.standard_normal(R.N)) if like == 1 else \
        rng.poisson(np.maximum(1 + dX_truth, 0)) * 1.0
    noise = 1.0 + 0.5 * rng.random(R.N)
    R.set_inputs(nobs=nobs, noise=noise)
    s = 0.5 * R.create_garfield(22, P)
    R.set_inputs(signal=s)
    mf, mr = R.hamiltonian_mass()
    mom = R.draw_momenta(23)
    g_ref = R.gradient_psi(s)
    pp_ref, pl_ref = R.psi(s)
    dX_ref = R.array("deltaX").copy()
    K_ref = R.kinetic(mom)
    traj = N <= 256                                     # (the 512^3 case stops here: minutes of CPU per evaluation)
    if traj:
        sf_ref, pf_ref = R.EoM(s, mom, 0.3, 1e-5)      # Neps = floor(8 * 0.3) + 1 = 3, eps = 1e-5
        neps, eps = int(R.scalar("Neps")), R.scalar("epsilon")
        dH_ref, sc_ref = R.delta_hamiltonian(s, mom, sf_ref, pf_ref)
    R.close()
    with Chain(Params(N1=N, L1=L, masskernel=mk, likelihood=like, rsd_model=rsd, calc_h=calc_h, mass_type=mass_type,
                      sfmodel=sfmodel, slength=cfg.slength)) as ch:
        ch.set_static(Power=P, nobs=nobs, noise=noise, window=one)
        mf_g, mr_g = ch.hamiltonian_mass()
        assert np.array_equal(mf_g.ravel(), mf) and np.array_equal(mr_g.ravel(), mr)
        assert rel_l2(ch.gradient_psi(s), g_ref) < TOL
        pp, pl, dX = ch.psi(s)
        assert abs(pp - pp_ref) <= TOL * abs(pp_ref) and abs(pl - pl_ref) <= TOL * abs(pl_ref)
        assert rel_l2(dX, dX_ref) < TOL
        assert abs(ch.kinetic_term(mom) - K_ref) <= TOL * abs(K_ref)
        if not traj:
            return
        sf, pf = ch.leapfrog(s, mom, neps, eps)
        assert rel_l2(sf, sf_ref) < 1e-8 and rel_l2(pf, pf_ref) < 1e-8
        dH, sc, _ = ch.delta_hamiltonian(s, mom, sf, pf)
        floor = 1e-13 * (abs(sc_ref["H_kin_i"]) + abs(sc_ref["psi_prior_i"]) + abs(sc_ref["psi_likeli_i"]))
        assert abs(dH - dH_ref) <= 1e-8 * abs(dH_ref) + floor


def test_device_resident_candidate_equals_the_separate_calls():
    """bgpu_candidate (momenta, trajectory and energies without leaving the device) against the same steps through
    the host-array entry points: draw_momenta_device -> kinetic / psi -> leapfrog -> kinetic / psi."""
    c = load_case("za_cic_gauss_rsd")
    with loaded_chain(c) as ch:
        s = c["signal"]
        neps, eps = 3, 1e-5
        mom = ch.draw_momenta_device(9, 4)
        Ki = ch.kinetic_term(mom)
        ppi, pli, _ = ch.psi(s)
        sf, pf = ch.leapfrog(s, mom, neps, eps)
        Kf = ch.kinetic_term(pf)
        ppf, plf, dXf = ch.psi(sf)
        ch.set_signal(s)
        E = ch.candidate(9, 4, neps, eps)
        for got, want in ((E["H_kin_i"], Ki), (E["psi_prior_i"], ppi), (E["psi_likeli_i"], pli), (E["H_kin_f"], Kf),
                          (E["psi_prior_f"], ppf), (E["psi_likeli_f"], plf)):
            assert abs(got - want) <= 1e-12 * abs(want)
        assert abs(E["momenta_f0"] - pf.ravel()[0]) <= 1e-13 * abs(pf.ravel()[0])
        x, dX = ch.accept()
        assert rel_l2(x, sf) < 1e-14 and rel_l2(dX, dXf) < 1e-13
        # the accepted field is the new current signal: the next candidate starts from it
        E2 = ch.candidate(9, 5, 1, eps)
        assert abs(E2["psi_prior_i"] - ppf) <= 1e-12 * abs(ppf)
        assert abs(E2["psi_likeli_i"] - plf) <= 1e-12 * abs(plf)
        # not accepted: the current signal and its energies stay (the library keeps psi of the current signal
        # instead of recomputing it per candidate as HMC.cc:214-215 does); a new signal drops them
        E3 = ch.candidate(9, 6, 1, eps)
        assert E3["psi_prior_i"] == E2["psi_prior_i"] and E3["psi_likeli_i"] == E2["psi_likeli_i"]
        ch.set_signal(s)
        E4 = ch.candidate(9, 7, 1, eps)
        assert abs(E4["psi_prior_i"] - ppi) <= 1e-12 * abs(ppi) and abs(E4["psi_likeli_i"] - pli) <= 1e-12 * abs(pli)


# ---------------------------------------------------------------- F4: mock data and initial guess on the device
@pytest.mark.parametrize("like,rho_c", [(1, 1.0), (0, 1.0), (0, 40.0), (3, 1.0)])
def test_mock_data_on_the_device(like, rho_c):
    """bgpu_mock_data = setup_random_test (barcoderunner.cc:42-205) with counter-based device streams: the truth has
    the spectrum it was drawn from, delta_eul is the chain's own forward model of it, and the observations follow
    their data model (Gaussian: unit-variance residuals; Poisson: counts with mean and variance Lambda, through both
    the inversion and the rejection branch of the sampler); the same seed reproduces the draw."""
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    N = 64
    L = inputs.box_length(N)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    with Chain(Params(N1=N, L1=L, masskernel=1, likelihood=like, rho_c=rho_c)) as ch:
        ch.set_static(Power=P)
        m = ch.mock_data(7, window_type=1, data_model=0, sigma_min=0.5, sigma_fac=0.1, negative_obs=True)
        m2 = ch.mock_data(7, window_type=1, data_model=0, sigma_min=0.5, sigma_fac=0.1, negative_obs=True)
        assert np.array_equal(m["delta_lag"], m2["delta_lag"])
        # (delta_eul comes out of unordered atomics: equal to rounding, and so are the observations drawn around it)
        if like == 0:
            assert np.mean(m["nobs"] != m2["nobs"]) < 1e-4
        else:
            assert rel_l2(m["nobs"], m2["nobs"]) < 1e-12
        m3 = ch.mock_data(8, window_type=10, data_model=0, sigma_min=0.5, sigma_fac=0.1, negative_obs=False)
        assert not np.array_equal(m["delta_lag"], m3["delta_lag"])
        assert np.all(m3["window"].ravel()[: n // 2] == 0) and np.all(m3["window"].ravel()[n // 2:] == 1)
        assert np.all(m3["nobs"].ravel()[: n // 2] == 0)
        assert like == 3 or np.all(m3["nobs"] >= 0)   # counts, or the Gaussian draw clamped at 0 (negative_obs off)
        # the truth's spectrum: <|d^|^2> = N^2/V P per mode (random.cpp:81-83), averaged over many modes
        dh = np.abs(np.fft.rfftn(m["delta_lag"])) ** 2
        Ph = P.reshape(N, N, N)[:, :, : N // 2 + 1]
        ok = Ph > 0
        ratio = dh[ok] / (float(n) ** 2 / L ** 3 * Ph[ok])
        assert abs(ratio.mean() - 1) < 0.02
        if like != 3:
            assert rel_l2(m["delta_eul"], ch.forward(m["delta_lag"])) < 1e-12
            assert abs(m["delta_eul"].mean()) < 1e-10
        lam = rho_c * (1 + m["delta_eul"]).ravel()
        nobs = m["nobs"].ravel()
        if like == 1:
            sig = 0.5 + 0.1 * lam
            assert rel_l2(m["noise"].ravel(), sig) < 1e-14
            z = (nobs - lam) / sig
            assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n)
        elif like == 0:
            assert np.all(nobs == np.round(nobs)) and np.all(nobs >= 0)
            lo, hi = lam < 30, lam >= 30            # CDF inversion / transformed rejection
            for sel in (lo, hi):
                if sel.sum() < 1000:
                    continue
                r = nobs[sel] - lam[sel]
                assert abs(r.mean()) < 5 * np.sqrt(lam[sel].mean() / sel.sum())
                assert abs(r.var() / lam[sel].mean() - 1) < 0.05
            if rho_c > 1:
                assert hi.sum() > 1000 and lo.sum() > 1000
        else:
            dl = m["delta_lag"].ravel()
            sig = 0.5 + 0.1 * dl * dl
            z = (nobs - dl) / sig
            assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n)


def test_initial_guess_on_the_device():
    """bgpu_initial_guess = make_initial_guess (barcoderunner.cc:207-247): zeros, a GRF with the prior's spectrum, the
    same smoothed with the Gaussian filter, or 0.1 * white noise."""
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    N = 64
    L = inputs.box_length(N)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    with Chain(Params(N1=N, L1=L)) as ch:
        ch.set_static(Power=P)
        assert np.all(ch.initial_guess(1, 0) == 0)
        w = ch.initial_guess(1, 4)
        assert abs(w.std() - 0.1) < 0.1 * 5 / np.sqrt(2 * n) and abs(w.mean()) < 0.1 * 5 / np.sqrt(n)
        g = ch.initial_guess(3, 2)
        gs = ch.initial_guess(3, 3, smoothing_scale=10.0)
        kx, ky, kz = (inputs.calc_ki(N, L)[:, None, None], inputs.calc_ki(N, L)[None, :, None],
                      inputs.calc_ki(N, L)[None, None, : N // 2 + 1])
        K = np.exp(-(kx ** 2 + ky ** 2 + kz ** 2) * 10.0 ** 2 / 2)
        assert rel_l2(np.fft.rfftn(gs), K * np.fft.rfftn(g)) < 1e-12   # same stream, filtered in k-space
        Ph = P.reshape(N, N, N)[:, :, : N // 2 + 1]
        ok = Ph > 0
        ratio = (np.abs(np.fft.rfftn(g)) ** 2)[ok] / (float(n) ** 2 / L ** 3 * Ph[ok])
        assert abs(ratio.mean() - 1) < 0.02
        with pytest.raises(Exception):
            ch.initial_guess(1, 1)

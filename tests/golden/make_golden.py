"""Generate the golden fixtures in tests/golden/ from the compiled reference.

Run in the build container (needs /root/reference and `make -C oracle`):

    python tests/golden/make_golden.py

Every output array below is produced by the UNMODIFIED reference sources
(oracle/_ref/libbarcode_ref.so via oracle/ref.py); nothing comes from the numpy
restatement or the CUDA path.  The reference's own tests hold no golden vector
for the hot path, so these files are the pin.

  ../../barcode_b200/data/pk_table.npz  the tabulated linear P(k) the reference reads (data/WMAP7_CAMB.dat),
                      in the float32 precision of calc_power.cc:41-42
  case_<name>.npz     inputs + reference outputs of one configuration at 16^3
  garfield_n8.npz     white-noise stream, coloured field and momenta at 8^3
  spectrum_n16.npz    measure_spectrum of a coloured field at 16^3
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402

CAMB = "/root/reference/data/WMAP7_CAMB.dat"

CASES = {
    # BASELINE.json configs[0]: ZA + CIC, Gaussian, real space
    "za_cic_gauss": dict(masskernel=1, likelihood=1, rsd_model=False, calc_h=0, mass_type=1),
    # configs[1]: (2LPT requested ->) ZA + CIC, Gaussian, RSD (HMC_models.cc:395-400)
    "za_cic_gauss_rsd": dict(masskernel=1, likelihood=1, rsd_model=True, calc_h=0, mass_type=1, sfmodel=2),
    # configs[2]: ZA + TSC, Poisson
    "za_tsc_poisson": dict(masskernel=2, likelihood=0, rsd_model=False, calc_h=0, mass_type=1),
    "za_ngp_gauss_h1": dict(masskernel=0, likelihood=1, rsd_model=False, calc_h=1, mass_type=1),
    "za_tsc_gauss_rsd_mass0": dict(masskernel=2, likelihood=1, rsd_model=True, calc_h=0, mass_type=0),
    "za_cic_poisson_mass4_dq": dict(masskernel=1, likelihood=0, rsd_model=False, calc_h=0, mass_type=4,
                                    deltaQ_factor=0.9, mass_factor=2.0),
    # Lag2Eul_non_zeldovich (2LPT + spherical collapse, ALPT split at slength; Lag2Eul.cc:138-312): every
    # sfmodel != 1 without rsd_model.  configs[3]'s forward model.
    "alpt_cic_gauss": dict(masskernel=1, likelihood=1, rsd_model=False, calc_h=0, mass_type=1, sfmodel=2,
                           slength=4.0),
    # the reference's shipped default (data/input.par:11-13,134): SPH spline mass assignment with its exact
    # adjoint, calc_h = 2 (likelihood_calc_h_SPH, HMC_models.cc:312-372)
    "za_sph_gauss_h2": dict(masskernel=3, likelihood=1, rsd_model=False, calc_h=2, mass_type=1),
    "za_sph_gauss_rsd_h2": dict(masskernel=3, likelihood=1, rsd_model=True, calc_h=2, mass_type=1,
                                particle_kernel_h_rel=1.3),
    "za_sph_poisson_h2": dict(masskernel=3, likelihood=0, rsd_model=False, calc_h=2, mass_type=4),
    # calc_h = 3: the Fourier / TSC variant of the SPH adjoint (likelihood_calc_V_SPH_fourier_TSC,
    # HMC_models_testing.cpp:54-188), interpolate_TSC's dz slip included
    "za_sph_gauss_h3": dict(masskernel=3, likelihood=1, rsd_model=False, calc_h=3, mass_type=1),
    "za_sph_gauss_rsd_h3": dict(masskernel=3, likelihood=1, rsd_model=True, calc_h=3, mass_type=1,
                                particle_kernel_h_rel=1.3),
    # log-normal likelihood (lognormal_independent.cpp): h = r, and the gradfindif product form with RSD in the
    # gradient's forward model but not in log_like's; a gentle signal keeps every cell occupied (the reference takes
    # the log of the unclamped density in the residual)
    "za_cic_lognormal_h1": dict(masskernel=1, likelihood=2, rsd_model=False, calc_h=1, mass_type=1, signal_scale=0.1,
                                eps_fac=1e-4),
    "za_tsc_lognormal_rsd": dict(masskernel=2, likelihood=2, rsd_model=True, calc_h=0, mass_type=1, signal_scale=0.1,
                                 delta_min=-0.9, eps_fac=1e-4),
    # Gaussian random field "likelihood" (gaussian_random_field.cpp, HMC.cc:159-160): no structure formation
    "grf": dict(masskernel=1, likelihood=3, rsd_model=False, calc_h=0, mass_type=1),
    # likelihood-force Hamiltonian masses (HMC_mass.cc:39-160): 1/P + force spectrum (2), + its mean (3)
    "za_cic_gauss_mass2": dict(masskernel=1, likelihood=1, rsd_model=False, calc_h=0, mass_type=2, N_bin=20),
    "za_tsc_gauss_rsd_mass3": dict(masskernel=2, likelihood=1, rsd_model=True, calc_h=0, mass_type=3, N_bin=20),
    "alpt_tsc_poisson_h1": dict(masskernel=2, likelihood=0, rsd_model=False, calc_h=1, mass_type=1, sfmodel=3,
                                slength=6.0, deltaQ_factor=0.95),
}

N1, L1 = 16, 50.0
NEPS_U, EPS_U = 0.3, 0.01   # N_eps_fac = 8 -> Neps = floor(8*0.3)+1 = 3 ; eps_fac = 1 -> eps = 0.01


def main():
    tab = np.loadtxt(CAMB)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(HERE)), "barcode_b200", "data", "pk_table.npz"), k=tab[:, 0].astype(np.float32),
                        P=tab[:, 1].astype(np.float32))

    only = set(sys.argv[1:])
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        kw = dict(kw)
        signal_scale = kw.pop("signal_scale", 0.5)
        eps_fac = kw.pop("eps_fac", 1.0)   # eps = eps_fac * EPS_U
        cfg = ref.Config(N1=N1, L1=L1, N_eps_fac=8.0, eps_fac=eps_fac, **kw)
        R = ref.Reference(cfg)
        P = R.readtab(CAMB)
        if kw.get("sfmodel", 1) != 1 and not kw.get("rsd_model", False):
            # the ALPT kernel file is re-read from the working directory on every forward evaluation
            import tempfile
            os.chdir(tempfile.mkdtemp(prefix="barcode_golden_"))
            R.kernelcomp()
        truth = R.create_garfield(1, P)
        one = np.ones(R.N)
        R.set_inputs(window=one, noise=one)
        dX_truth = R.forward(truth, want_pos=False)
        rng = np.random.default_rng(11)
        if cfg.likelihood == 1:
            nobs = np.maximum(0.0, 1.0 + dX_truth + rng.standard_normal(R.N))
        elif cfg.likelihood == 2:   # barcoderunner.cc:163-183: log of the clamped density plus noise
            nobs = np.log(1.0 + np.maximum(dX_truth, cfg.delta_min)) + 0.3 * rng.standard_normal(R.N)
        elif cfg.likelihood == 3:   # the data ARE a noisy Lagrangian field
            nobs = truth + rng.standard_normal(R.N)
        else:
            nobs = rng.poisson(np.maximum(1.0 + dX_truth, 0.0)).astype(np.float64)
        # a window with a masked corner exercises the w > 0 branches
        window = one.copy().reshape(N1, N1, N1)
        window[:3, :3, :3] = 0.0
        window = window.ravel()
        noise = 1.0 + 0.25 * rng.random(R.N)
        R.set_inputs(nobs=nobs, window=window, noise=noise)
        s = signal_scale * R.create_garfield(2, P)
        R.set_inputs(signal=s)
        mass_f, mass_r = R.hamiltonian_mass()
        mom = R.draw_momenta(3)
        out = dict(Power=P, nobs=nobs, noise=noise, window=window, signal=s, momenta=mom,
                   mass_f=mass_f, mass_r=mass_r, truth=truth)
        dX, x, y, z = R.forward(s)
        out.update(deltaX_fwd=dX, posx=x, posy=y, posz=z)
        out["grad_like"] = R.grad_log_like(s)
        out["grad_prior"] = R.grad_log_prior(s)
        out["gradpsi"] = R.gradient_psi(s)
        pp, pl = R.psi(s)
        out["psi_prior"], out["psi_like"] = pp, pl
        out["deltaX_psi"] = R.array("deltaX").copy()
        out["K"] = R.kinetic(mom)
        sf, pf = R.EoM(s, mom, NEPS_U, EPS_U)
        out["Neps"], out["epsilon"] = R.scalar("Neps"), R.scalar("epsilon")
        out["s_f"], out["p_f"] = sf, pf
        dH, sc = R.delta_hamiltonian(s, mom, sf, pf)
        out["dH"] = dH
        for k, v in sc.items():
            out["dh_" + k] = v
        out["D1"] = R.scalar("D1")
        out["D2"] = R.scalar("D2")
        out["cfg"] = np.array(repr({**kw, "N1": N1, "L1": L1}))
        out["signal_scale"] = signal_scale
        out["eps_fac"] = eps_fac
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
        print(name, "gradpsi norm", np.linalg.norm(out["gradpsi"]), "dH", dH, "Neps", out["Neps"])
        R.close()

    if only and "spectrum" not in only:
        return
    # measure_spectrum (field_statistics.cpp:20-90) of a coloured field at 16^3, 20 and 200 bins (the latter
    # leaves bins empty)
    R = ref.Reference(ref.Config(N1=N1, L1=L1))
    P = R.readtab(CAMB)
    field = R.create_garfield(5, P)
    R.close()
    out = dict(signal=field, L1=L1)
    for nb in (20, 200):
        km, pw = ref.measure_spectrum(field.reshape(N1, N1, N1), L1, nb)
        out[f"kmode_{nb}"], out[f"power_{nb}"] = km, pw
    np.savez_compressed(os.path.join(HERE, "spectrum_n16.npz"), **out)
    if only:
        return
    # momentum draw at 8^3: the white-noise stream, its colouring with P and with 1/P
    cfg = ref.Config(N1=8, L1=25.0)
    R = ref.Reference(cfg)
    P = R.readtab(CAMB)
    white = ref.white_noise(8, 7)
    field = R.create_garfield(7, P)
    R.set_inputs(window=np.ones(R.N), noise=np.ones(R.N), nobs=np.ones(R.N))
    mass_f, _ = R.hamiltonian_mass()
    mom = R.draw_momenta(7)
    raw, gauss = ref.rng_stream(7, 16, 16)
    np.savez_compressed(os.path.join(HERE, "garfield_n8.npz"), Power=P, white=white, field=field, mass_f=mass_f,
                        momenta=mom, raw=raw, gauss=gauss)
    R.close()


if __name__ == "__main__":
    main()

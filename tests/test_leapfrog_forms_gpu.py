"""The three forms of Hamiltonian_EoM on the device (api.cu leapfrog_device) walk the same trajectory.

  * step by step, as the reference writes it (HMC.cc:251-369: two half kicks and a host test of momenta[0] per step;
    BGPU_LEAPFROG_FUSED=0),
  * fused in real space (merged kicks in the gradient's last store, device run-away flag; BGPU_LEAPFROG_KSPACE=0),
  * in k-space (default where it applies: Fourier-space mass, Zel'dovich model).
The golden trajectories (tests/test_gpu_parity.py::test_leapfrog_and_delta_H) pin whichever form is the default to
the compiled reference; this file pins the forms to one another, including a trajectory the run-away test stops.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

FORMS = {"reference": {"BGPU_LEAPFROG_FUSED": "0"}, "fused": {"BGPU_LEAPFROG_KSPACE": "0"}, "kspace": {}}


def run(monkeypatch, form, kw, prob, mom, Neps, eps, mass_f=None):
    from barcode_b200.chain import Chain, Params
    for k in ("BGPU_LEAPFROG_FUSED", "BGPU_LEAPFROG_KSPACE"):
        monkeypatch.delenv(k, raising=False)
    for k, v in FORMS[form].items():
        monkeypatch.setenv(k, v)   # read by bgpu_create
    with Chain(Params(**kw)) as ch:
        ch.set_static(Power=prob["Power"], nobs=prob["nobs"], noise=prob["noise"], window=prob["window"])
        if mass_f is None:
            ch.hamiltonian_mass()
        else:
            ch.set_mass(mass_f=mass_f)
        return ch.leapfrog(prob["signal"], mom, Neps, eps)


def problem(kw, seed=5):
    from barcode_b200.chain import Chain, Params
    from barcode_b200 import inputs
    with Chain(Params(**kw)) as ch:
        return inputs.synthetic_problem(ch, seed=seed)


@pytest.mark.parametrize("like,rsd,calc_h,masskernel", [(1, True, 0, 1), (1, True, 4, 1), (0, False, 0, 2), (1, False, 4, 2)])
def test_the_three_forms_walk_the_same_trajectory(monkeypatch, like, rsd, calc_h, masskernel):
    from barcode_b200 import inputs
    N = 32
    kw = dict(N1=N, L1=inputs.box_length(N), masskernel=masskernel, likelihood=like, rsd_model=rsd, calc_h=calc_h,
              mass_type=1, sfmodel=1)
    prob = problem(kw)
    out = {f: run(monkeypatch, f, kw, prob, prob["momenta"], 5, 2e-3) for f in FORMS}
    moved = rel_l2(out["reference"][0], prob["signal"])
    assert moved > 1e-6, "the trajectory must go somewhere for the comparison to mean anything"
    # Rounding only.  The forms sum the same terms in different orders, and the scatter's reductions land in a
    # different order every run.  Measured on B200 over repeated runs: Gaussian 4e-15 ... 1e-14, Poisson (whose
    # residual 1 - n / Lambda amplifies the density's rounding in nearly empty cells) 1e-13 ... 4e-12, exact TSC
    # adjoint 1e-14; the exact CIC adjoint (calc_h = 4 with masskernel 1) differentiates a piecewise-LINEAR weight
    # whose slope jumps at cell faces, so a particle that two roundings of the same s put on different sides of a
    # face changes its gradient by O(1): 2e-10 in s_f and 6e-9 ... 2e-8 in p_f, between two runs of the SAME form
    # as much as between forms.  A wrong factor or a missing update would show at 1e-3.
    tol = 1e-6 if (calc_h == 4 and masskernel == 1) else (1e-9 if like == 0 else 1e-11)
    errs = {f: (rel_l2(out[f][0], out["reference"][0]), rel_l2(out[f][1], out["reference"][1])) for f in ("fused", "kspace")}
    print("leapfrog forms", like, rsd, calc_h, masskernel, {f: "%.2e %.2e" % e for f, e in errs.items()}, "tol %.0e" % tol)
    for f, e in errs.items():
        assert e[0] < tol and e[1] < tol, (f, e)


def test_a_runaway_trajectory_stops_after_the_same_step_in_every_form(monkeypatch):
    """HMC.cc:360-364: |momenta[0]| > 1e50 after a step ends the trajectory.  Momenta of 3e50 with a mass that keeps the
    drift gentle: every form must return the state after ONE step, whatever Neps asks for."""
    from barcode_b200 import inputs
    N = 32
    kw = dict(N1=N, L1=inputs.box_length(N), masskernel=1, likelihood=1, rsd_model=True, calc_h=0, mass_type=1, sfmodel=1)
    prob = problem(kw)
    rng = np.random.default_rng(11)
    mom = 3e50 * (1.0 + 0.1 * rng.standard_normal((N, N, N)))
    mass = np.full((N, N, N), 3e53)
    one = run(monkeypatch, "reference", kw, prob, mom, 1, 1e-3, mass_f=mass)
    three_free = run(monkeypatch, "reference", kw, prob, 1e-51 * mom, 3, 1e-3, mass_f=1e-51 * mass)   # no run-away
    assert rel_l2(one[0], prob["signal"]) > 1e-6
    # without the stop three steps drift three times as far as one
    assert abs(rel_l2(three_free[0], prob["signal"]) / rel_l2(one[0], prob["signal"]) - 3.0) < 0.2
    for f in FORMS:
        sf, pf = run(monkeypatch, f, kw, prob, mom, 3, 1e-3, mass_f=mass)
        assert rel_l2(sf, one[0]) < 1e-10, f
        assert rel_l2(pf, one[1]) < 1e-10, f

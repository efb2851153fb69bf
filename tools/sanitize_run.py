"""Small driver for compute-sanitizer (memcheck): one pass over every kernel family at 32^3 (first-generation
passes) and 128^3 (TMA-staged passes).

    python tools/sanitize_run.py            (compute-sanitizer --tool memcheck ... where the pool allows it; closed on this one)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barcode_b200 import chain as bc, inputs  # noqa: E402

for N in (32, 128):
    L = inputs.box_length(N)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    rng = np.random.default_rng(N)
    n = N ** 3
    s = 0.3 * rng.standard_normal(n)
    nobs = np.maximum(0.0, 1.0 + 0.2 * rng.standard_normal(n))
    for kw in (dict(masskernel=1, likelihood=1, rsd_model=True, calc_h=0, mass_type=1),
               dict(masskernel=2, likelihood=0, rsd_model=False, calc_h=4, mass_type=0),
               dict(masskernel=3, likelihood=1, rsd_model=False, calc_h=2, mass_type=4),
               dict(masskernel=1, likelihood=2, rsd_model=False, calc_h=0, mass_type=1),
               dict(masskernel=1, likelihood=1, rsd_model=False, calc_h=4, mass_type=2, sfmodel=2, slength=8.0, N_bin=16),
               dict(masskernel=0, likelihood=3, rsd_model=False, calc_h=1, mass_type=1)):
        with bc.Chain(bc.Params(N1=N, L1=L, **kw)) as ch:
            ch.set_static(Power=P, nobs=nobs, noise=np.ones(n), window=np.ones(n))
            ch.hamiltonian_mass(s)
            g = ch.gradient_psi(s)
            pp, pl, dX = ch.psi(s)
            mom = ch.draw_momenta_device(1, 0)
            K = ch.kinetic_term(mom)
            sf, pf = ch.leapfrog(s, mom, 1, 1e-4)
            km, pw = ch.measure_spectrum(s, 16)
            ok = np.isfinite(g).all() and np.isfinite([pp, pl, K]).all() and np.isfinite(sf).all()
            print(N, kw, "finite" if ok else "NON-FINITE", flush=True)
print("sanitize_run done")

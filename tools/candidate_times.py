"""Wall-clock time of the device-resident candidate and of its pieces (developer tool).

    python tools/candidate_times.py --grid 64 [--neps 2]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barcode_b200 import chain as bc, inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--neps", type=int, default=2)
a = ap.parse_args()
N = a.grid
ch = bc.Chain(bc.Params(N1=N, L1=inputs.box_length(N), masskernel=1, likelihood=1, rsd_model=True, sfmodel=1, calc_h=0, mass_type=1))
prob = inputs.synthetic_problem(ch, seed=2)
s = prob["signal"]
ch.set_signal(s)


def t(f, n=20):
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return 1e3 * (time.perf_counter() - t0) / n


i = [0]
def cand():
    i[0] += 1
    ch.candidate(9, i[0], a.neps, 1e-4)

print(f"grid {N}: candidate (Neps = {a.neps}) {t(cand):.3f} ms")
mom = ch.draw_momenta_device(9, 4)
print(f"  draw_momenta_device (host out) {t(lambda: ch.draw_momenta_device(9, 5)):.3f} ms")
print(f"  kinetic (host in)              {t(lambda: ch.kinetic_term(mom)):.3f} ms")
print(f"  psi (host in)                  {t(lambda: ch.psi(s)):.3f} ms")
print(f"  leapfrog (host in / out)       {t(lambda: ch.leapfrog(s, mom, a.neps, 1e-4)):.3f} ms")
print(f"  gradient_psi (host in / out)   {t(lambda: ch.gradient_psi(s)):.3f} ms")

"""Developer check of the TMA-staged strided pass against numpy and against the cp.async pass.

    python tools/fft_check.py [--grids 128 256]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barcode_b200 import chain as bc, inputs  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


ap = argparse.ArgumentParser()
ap.add_argument("--grids", type=int, nargs="+", default=[128, 256])
ap.add_argument("--calc-h", type=int, nargs="+", default=[0, 4])
a = ap.parse_args()
ok = True
for N in a.grids:
    rng = np.random.default_rng(N)
    x = rng.standard_normal((N, N, N))
    ref = np.fft.rfftn(x)
    res = {}
    for tma in ("1", "0"):
        os.environ["BGPU_FFT_TMA"] = tma
        with bc.Chain(bc.Params(N1=N, L1=inputs.box_length(N), masskernel=1, likelihood=1, rsd_model=True,
                                calc_h=0)) as ch:
            c = ch.fft_r2c(x)
            back = ch.fft_c2r(ref)
            e1, e2 = rel(c, ref), rel(back, x)
            print(f"N={N} tma={tma}: r2c err {e1:.2e}  c2r err {e2:.2e}", flush=True)
            ok &= e1 < 1e-14 and e2 < 1e-14
            P = inputs.power_on_grid(*inputs.load_pk_table(), N, inputs.box_length(N))
            cv = ch.convolve_inv_corr(x, P)
            res[("conv", tma)] = cv
        for calc_h in a.calc_h:
            with bc.Chain(bc.Params(N1=N, L1=inputs.box_length(N), masskernel=1, likelihood=1, rsd_model=True,
                                    calc_h=calc_h)) as ch:
                n = N ** 3
                ones = np.ones(n)
                ch.set_static(Power=P, nobs=ones + 0.1 * rng.standard_normal(n), noise=ones, window=ones)
                rng2 = np.random.default_rng(7)
                s = 0.3 * rng2.standard_normal(n)
                g = ch.gradient_psi(s)
                res[("grad", calc_h, tma)] = g
                t0 = time.time()
                for _ in range(3):
                    g = ch.gradient_psi(s)
                print(f"   calc_h={calc_h}: {(time.time() - t0) / 3 * 1e3:.2f} ms per host-call gradient", flush=True)
                rng = np.random.default_rng(N)
                rng.standard_normal((N, N, N))
    e = rel(res[("conv", "1")], res[("conv", "0")])
    print(f"N={N}: convolve tma vs cp.async {e:.2e}")
    ok &= e < 1e-13
    for calc_h in a.calc_h:
        e = rel(res[("grad", calc_h, "1")], res[("grad", calc_h, "0")])
        print(f"N={N}: gradient calc_h={calc_h} tma vs cp.async {e:.2e}")
        ok &= e < 1e-12
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)

"""Per-kernel-class device times of one gradient evaluation (developer tool).

    python tools/kernel_times.py --grid 512 [--calc-h 0] [--steps 3]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barcode_b200 import chain as bc, inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--calc-h", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
N = a.grid
n = N ** 3
nh = N * N * (N // 2 + 1)
ch = bc.Chain(bc.Params(N1=N, L1=inputs.box_length(N), masskernel=1, likelihood=1, rsd_model=True, sfmodel=2,
                        calc_h=a.calc_h))
rng = np.random.default_rng(0)
P = inputs.power_on_grid(*inputs.load_pk_table(), N, inputs.box_length(N))
ones = np.ones(n)
ch.set_static(Power=P, nobs=ones + 0.1 * rng.standard_normal(n), noise=ones, window=ones)
s = 0.3 * rng.standard_normal(n)
for _ in range(2):
    g = ch.gradient_psi(s)
bc.profile_begin()
for _ in range(a.steps):
    g = ch.gradient_psi(s)
prof = bc.profile_end()
alg = {"fft_strided_pass_y": 2 * nh * 16, "fft_strided_pass_x": 2 * nh * 16, "fft_r2c_zpass": n * 8 + nh * 16, "fft_c2r_zpass": n * 8 + nh * 16,
       "scatter": 4 * n * 8, "gather_adjoint": 7 * n * 8, "overdens_residual": 5 * n * 8, "reduce": n * 8,
       "stream": 3 * n * 8, "colour_momenta": n * 16 + nh * 16}
tot = sum(v[0] for v in prof.values())
print(f"grid {N} calc_h {a.calc_h}: {tot / a.steps:.3f} ms per gradient (sum of kernels), {1e3 * a.steps / tot:.1f} evals/s")
for k, (ms, cnt) in prof.items():
    if cnt:
        print(f"  {k:20s} {ms / a.steps:8.3f} ms/step  {cnt / a.steps:5.1f} launches  {ms / cnt * 1e3:8.1f} us each  "
              f"{alg[k] / (ms / cnt * 1e-3) / 1e9:7.0f} GB/s")

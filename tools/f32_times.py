"""Per-kernel-class device times of one single-precision gradient evaluation (developer tool; ncu target).

    python tools/f32_times.py --grid 256 [--calc-h 0] [--steps 3]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barcode_b200 import chain as bc, inputs  # noqa: E402
from barcode_b200.chain_f32 import ChainF32  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--calc-h", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
N = a.grid
n = N ** 3
ch = ChainF32(bc.Params(N1=N, L1=inputs.box_length(N), masskernel=1, likelihood=1, rsd_model=True, sfmodel=1,
                        calc_h=a.calc_h))
rng = np.random.default_rng(0)
P = inputs.power_on_grid(*inputs.load_pk_table(), N, inputs.box_length(N))
ones = np.ones(n, dtype=np.float32)
ch.set_static(Power=P, nobs=ones + 0.1 * rng.standard_normal(n).astype(np.float32), noise=ones, window=ones)
s = (0.3 * rng.standard_normal(n)).astype(np.float32)
for _ in range(2):
    g = ch.gradient_psi(s)
bc.profile_begin()
for _ in range(a.steps):
    g = ch.gradient_psi(s)
prof = bc.profile_end()
tot = sum(v[0] for v in prof.values())
print(f"f32 grid {N} calc_h {a.calc_h}: {tot / a.steps:.3f} ms per gradient (sum of kernels)")
for k, (ms, cnt) in prof.items():
    if cnt:
        print(f"  {k:20s} {ms / a.steps:8.3f} ms/step  {cnt / a.steps:5.1f} launches  {ms / cnt * 1e3:8.1f} us each")

"""Host-side helpers of the package (no GPU): P(k) table, k vectors, bench plumbing."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_calc_ki_convention():
    from barcode_b200 import inputs
    k = inputs.calc_ki(8, 2 * np.pi)        # kfac = 1
    assert np.array_equal(k, [0, 1, 2, 3, 4, -3, -2, -1])   # Nyquist positive (scale_space.cpp:41-51)


def test_pk_table_and_grid():
    from barcode_b200 import inputs
    k, P = inputs.load_pk_table()
    assert k.dtype == np.float32 and len(k) == len(P) > 700 and k[0] == 0.0
    G = inputs.power_on_grid(k, P, 16, 50.0)
    assert G[0, 0, 0] == 0.0 and np.all(G.ravel()[1:] > 0)
    # |k| symmetry of the full real-indexed grid
    assert G[1, 2, 3] == G[15, 14, 13] == G[1, 14, 3]
    # chunked evaluation is chunk-size independent
    assert np.array_equal(G, inputs.power_on_grid(k, P, 16, 50.0, chunk=5))


def test_white_noise_statistics():
    from barcode_b200 import inputs
    W = inputs.complex_white_noise(16, 3)
    assert W.shape == (16, 16, 16) and abs(W.real.std() - 1) < 0.05 and abs(W.imag.std() - 1) < 0.05


def test_bench_reference_arm_small_grid():
    """`bench.py --impl reference` end to end on CPU at a tiny grid (the driver runs it at 256^3)."""
    from oracle import ref
    if not ref.available():
        import pytest
        pytest.skip("oracle/_ref not built")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid",
                                   "16", "--steps", "2", "--warmup", "1"], cwd=ROOT)
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "hmc_gradient_evals_per_s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0

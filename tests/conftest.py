import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASE_NAMES = ["za_cic_gauss", "za_cic_gauss_rsd", "za_tsc_poisson", "za_ngp_gauss_h1", "za_tsc_gauss_rsd_mass0",
              "za_cic_poisson_mass4_dq", "alpt_cic_gauss", "alpt_tsc_poisson_h1",
              "za_sph_gauss_h2", "za_sph_gauss_rsd_h2", "za_sph_poisson_h2", "za_sph_gauss_h3", "za_sph_gauss_rsd_h3",
              "za_cic_lognormal_h1", "za_tsc_lognormal_rsd", "grf", "za_cic_gauss_mass2", "za_tsc_gauss_rsd_mass3"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_case(name):
    with np.load(os.path.join(GOLDEN, f"case_{name}.npz")) as f:
        d = {k: f[k] for k in f.files}
    d["cfg"] = eval(str(d["cfg"]))  # written by tests/golden/make_golden.py
    return d


def rel_l2(a, b):
    a, b = np.ravel(a), np.ravel(b)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(params=CASE_NAMES)
def case(request):
    return load_case(request.param)

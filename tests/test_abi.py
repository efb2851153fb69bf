"""The C-ABI shared library: builds, loads, exports every symbol include/barcode_gpu.h
declares, and fails loudly (no CPU fallback) where there is no GPU.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "barcode_gpu.h")


@pytest.fixture(scope="module")
def lib():
    from barcode_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from barcode_b200 import build
        build.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_typed(lib):
    from barcode_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/barcode_gpu.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in barcode_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_abi_version_and_defaults(lib):
    from barcode_b200._lib import BgpuParams
    assert lib.bgpu_abi_version() == 2
    p = BgpuParams()
    lib.bgpu_default_params(C.byref(p))
    assert (p.N1, p.N2, p.N3) == (64, 64, 64) and p.L1 == 200.0          # data/input.par:117-125
    assert p.masskernel == 1 and p.likelihood == 1 and p.mass_type == 1
    assert p.rho_c == p.biasP == p.biasE == 1.0                             # init_par.cc:574-578
    assert p.OM == 0.272 and abs(p.OL - 0.728) < 1e-15                      # init_par.cc:38,480-483


def test_params_struct_layout_matches_header(lib):
    """sizeof(bgpu_params) as the C compiler sees it == the ctypes mirror."""
    import subprocess
    import tempfile
    from barcode_b200._lib import BgpuParams
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "barcode_gpu.h"\n'
                             'int main(){printf("%zu %zu %zu", sizeof(bgpu_params), offsetof(bgpu_params, D1), '
                             'offsetof(bgpu_params, device));return 0;}')
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        size, off_d1, off_dev = map(int, subprocess.check_output([exe]).split())
    assert size == C.sizeof(BgpuParams)
    assert off_d1 == BgpuParams.D1.offset and off_dev == BgpuParams.device.offset


def test_validation_errors_are_reported_not_swallowed(lib):
    from barcode_b200._lib import BgpuParams
    h = C.c_void_p()
    for field, val, needle in (("N1", 48, "power of two"), ("masskernel", 4, "masskernel"), ("calc_h", 2, "SPH"),
                               ("calc_h", 3, "calc_h"), ("likelihood", 4, "likelihood"),
                               ("mass_type", 5, "mass_type")):
        p = BgpuParams()
        lib.bgpu_default_params(C.byref(p))
        setattr(p, field, val)
        if field == "N1":
            p.N2 = p.N3 = val
        assert lib.bgpu_create(C.byref(p), C.byref(h)) != 0
        assert needle in lib.bgpu_last_error().decode()
        assert not h.value


def test_no_cpu_fallback_without_a_gpu(lib):
    """On a machine without a CUDA device creation must fail with a clear message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here; the no-device branch is exercised on the CPU box")
    from barcode_b200.chain import Chain, Params
    from barcode_b200._lib import BgpuError
    with pytest.raises(BgpuError, match="no CUDA device|CUDA"):
        Chain(Params(N1=16, L1=50.0))

"""The slab-decomposed chain on real GPUs (needs >= 2): tools/slab_check.py under torchrun compares it,
quantity by quantity, with the single-GPU chain (which the parity tests pin to the reference)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("p2p,generic,sfmodel,extra", [
    ("1", "0", 1, []), ("0", "0", 1, []), ("0", "1", 1, []), ("1", "0", 2, []),
    ("1", "0", 1, ["--likelihood", "0", "--mass-type", "2", "--masskernel", "2"]),
    ("1", "0", 1, ["--likelihood", "2", "--mass-type", "3", "--amp", "0.1"])])
def test_slab_chain_matches_single_gpu_chain(p2p, generic, sfmodel, extra):
    """p2p: fused transpose over peer memory vs NCCL all-to-all; generic: the cp.async slab pass that
    1024^3 runs on (fft_slab_generic.cuh), here at 128^3 where a single-GPU reference exists; sfmodel 2: the
    2LPT/ALPT forward model with its stencil and cell-boundary halos (BASELINE.json configs[3]); extra: Poisson +
    TSC / log-normal with the finite-difference product gradient (2-plane halo of delta_x) and the likelihood-force
    masses.  Every run also checks calc_h 0 / 1 / 4, measure_spectrum, the device momentum draw and a leapfrog."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, BGPU_SLAB_P2P=p2p, BGPU_FFT_SLAB_GENERIC=generic)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29655",
                        os.path.join(ROOT, "tools", "slab_check.py"), "--grid", "128", "--sfmodel", str(sfmodel)] + extra,
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + "\n...\n" + r.stderr[-6000:]
    assert "OK" in r.stdout

"""Size-independent checks of the slab-decomposed chain at grids too large for a single-GPU reference
(1024^3; run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
        tools/slab_props.py --grid 1024

  * plane waves: r2c of cos / sin modes puts N^3/2 into exactly the right (kx, ky, kz) bins of the
    transposed k-space slabs and ~0 everywhere else (pins every index of the distributed transform);
  * white noise: c2r(r2c(w)) = w, Parseval;
  * forward model: sum(rho) = N^3 (mean delta_x = 0), finite gradient, psi_likelihood >= 0.
Exit code 0 = all within tolerance.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barcode_b200 import chain as bc, inputs, multi, slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    info = multi.rank_info()
    torch.cuda.set_device(info.local_rank)
    multi.init("nccl", info, torch.device("cuda", info.local_rank))
    N = a.grid
    L = inputs.box_length(N)
    kw = dict(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=0, mass_type=1)
    sc = slab.SlabChain.create(bc.Params(device=info.local_rank, **kw), info.rank, info.world)
    ok = True
    msgs = []

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- plane waves
    x = (sc.x0 + np.arange(sc.Ns))[:, None, None]
    y = np.arange(N)[None, :, None]
    z = np.arange(N)[None, None, :]
    m1, m2 = (3, 5, 7), (N - 11, 2, 1)
    ph1 = (2 * np.pi / N) * ((m1[0] * x + m1[1] * y + m1[2] * z) % N)
    w = np.cos(ph1)
    del ph1
    ph2 = (2 * np.pi / N) * ((m2[0] * x + m2[1] * y + m2[2] * z) % N)
    w += 0.5 * np.sin(ph2)
    del ph2
    W = sc.fft_r2c(w).reshape(sc.kshape)           # [x][y_local][z]
    half = 0.5 * float(N) ** 3
    expect = {m1: half + 0j, m2: -0.25j * float(N) ** 3}
    err_peak = 0.0
    for (kx, ky, kz), val in expect.items():
        if sc.x0 <= ky < sc.x0 + sc.Ns:
            err_peak = max(err_peak, abs(W[kx, ky - sc.x0, kz] - val) / half)
            W[kx, ky - sc.x0, kz] = 0.0
    err_peak = allmax(err_peak)
    leak = allmax(np.abs(W).max() / half)
    msgs.append(f"plane waves: peak error {err_peak:.2e}, leakage {leak:.2e}")
    ok &= err_peak < 1e-12 and leak < 1e-12
    del W

    # ---- white noise round trip + Parseval
    w = np.random.default_rng(17 + info.rank).standard_normal(sc.shape)
    W = sc.fft_r2c(w).reshape(sc.kshape)
    wt = np.full(N // 2 + 1, 2.0)
    wt[0] = wt[-1] = 1.0
    p_k = allsum(float(np.einsum("xyz,z->", np.abs(W) ** 2, wt))) / float(N) ** 3
    p_x = allsum(float(np.sum(w * w)))
    back = sc.fft_c2r(W)
    rt = np.sqrt(allsum(float(np.sum((back.reshape(sc.shape) - w) ** 2))) / p_x)
    msgs.append(f"white noise: round trip {rt:.2e}, Parseval {abs(p_k / p_x - 1):.2e}")
    ok &= rt < 1e-14 and abs(p_k / p_x - 1) < 1e-13
    del W, back, w

    # ---- forward model + gradient on the synthetic problem of the bench
    t0 = time.time()
    prob = inputs.slab_problem(sc, seed=1)
    t_prob = time.time() - t0
    d = sc.forward(prob["signal"])
    mean_delta = allsum(float(d.sum())) / float(N) ** 3
    pp, pl, _ = sc.psi(prob["signal"])
    g = sc.gradient_psi(prob["signal"])
    gnorm = np.sqrt(allsum(float(np.sum(g * g))))
    finite = allsum(float(np.count_nonzero(~np.isfinite(g)))) == 0
    msgs.append(f"forward: mean delta {mean_delta:.2e}; psi prior {pp:.6e} likelihood {pl:.6e}; |grad| {gnorm:.6e} "
                f"finite={finite}; inputs built in {t_prob:.0f} s")
    ok &= abs(mean_delta) < 1e-12 and finite and pl >= 0 and gnorm > 0

    # ---- speed (device-resident)
    d_s = torch.from_numpy(np.ascontiguousarray(prob["signal"]).reshape(-1)).cuda()
    d_g = torch.empty_like(d_s)
    stream = torch.cuda.current_stream()
    sc.set_stream(stream.cuda_stream)
    sc.gradient_psi_dev(d_s.data_ptr(), d_g.data_ptr())
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        sc.gradient_psi_dev(d_s.data_ptr(), d_g.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = multi.max_over_ranks(e0.elapsed_time(e1), info, "cuda") / a.steps
    sc.close()
    if info.rank == 0:
        for m in msgs:
            print("  " + m)
        print(f"  gradient evaluation: {ms:.1f} ms ({1e3 / ms:.2f} evals/s) on {info.world} GPU(s)")
        print(f"slab_props grid {N} on {info.world} ranks:", "OK" if ok else "MISMATCH", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    multi.finalize()
    return int(flag.item() != 0)


if __name__ == "__main__":
    sys.exit(main())

"""Slab-decomposed chain vs the single-GPU chain on the same inputs (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/slab_check.py --grid 128

Rank 0 also runs the cube chain and compares: FFT round trip, convolve, forward density, psi, kinetic
energy, gradient (calc_h 0 / 1), a short leapfrog trajectory.  Exit code 0 = all within tolerance.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barcode_b200 import chain as bc, inputs, multi, slab  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--masskernel", type=int, default=1)
    ap.add_argument("--amp", type=float, default=0.5)
    ap.add_argument("--sfmodel", type=int, default=1, help="2/3: Lag2Eul_non_zeldovich (forces rsd_model off)")
    ap.add_argument("--likelihood", type=int, default=1)
    ap.add_argument("--mass-type", type=int, default=1)
    a = ap.parse_args()
    info = multi.rank_info()
    torch.cuda.set_device(info.local_rank)
    multi.init("nccl", info, torch.device("cuda", info.local_rank))
    N = a.grid
    L = inputs.box_length(N)
    rng = np.random.default_rng(3)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    nobs = np.maximum(0.0, 1.0 + 0.3 * rng.standard_normal(n)).reshape(N, N, N)
    noise = np.ones((N, N, N))
    window = np.ones((N, N, N))
    # a smooth-ish signal with displacements of a few cells: white noise coloured by sqrt(P)
    w = rng.standard_normal((N, N, N))
    s = np.fft.irfftn(np.fft.rfftn(w) * np.sqrt(np.maximum(P[:, :, :N // 2 + 1], 0) * n / L ** 3), s=(N, N, N), axes=(0, 1, 2))
    s *= a.amp
    # momenta with the spectrum of the mass 1/P, scaled so that a drift moves s by ~10 %
    Ph = np.maximum(P[:, :, :N // 2 + 1], 0)
    invPh = np.where(Ph > 0, 1.0 / np.where(Ph > 0, Ph, 1.0), 0.0)
    p0 = np.fft.irfftn(np.fft.rfftn(rng.standard_normal((N, N, N))) * np.sqrt(invPh), s=(N, N, N), axes=(0, 1, 2))
    drift = np.fft.irfftn(np.fft.rfftn(p0) * (L ** 3 / n) * Ph, s=(N, N, N), axes=(0, 1, 2))
    p0 *= 0.1 * np.abs(s).max() / (np.abs(drift).max() * 2e-3)
    ok = True
    worst = {}

    def gather(local):
        t = torch.from_numpy(np.ascontiguousarray(local)).cuda()
        out = [torch.empty_like(t) for _ in range(info.world)]
        dist.all_gather(out, t)
        return torch.cat(out, 0).cpu().numpy()

    for calc_h in (0, 1, 4):
        if calc_h == 4 and a.sfmodel != 1:
            continue   # the exact adjoint of the 2LPT/ALPT model is single-GPU only (bgpu_slab_create says so)
        kw = dict(N1=N, L1=L, masskernel=a.masskernel, likelihood=a.likelihood, rsd_model=(a.sfmodel == 1),
                  calc_h=calc_h, mass_type=a.mass_type, sfmodel=a.sfmodel, N_bin=40)
        sc = slab.SlabChain.create(bc.Params(device=info.local_rank, **kw), info.rank, info.world)
        sc.set_static(Power=sc.local(P), nobs=sc.local(nobs), noise=sc.local(noise), window=sc.local(window))
        res_mass = gather(sc.hamiltonian_mass(sc.local(s))[0]) if a.mass_type in (1, 2, 3, 4) else None
        res = {}
        if res_mass is not None:
            res["mass_f"] = res_mass
        if calc_h == 0:
            km, pw = sc.measure_spectrum(sc.local(s), 40)
            res["spectrum"] = np.concatenate([km, pw])
            res["fft_roundtrip"] = gather(sc.fft_c2r(sc.fft_r2c(sc.local(s))))
            res["convolve"] = gather(sc.convolve_inv_corr(sc.local(s), sc.local(P)))
            res["forward"] = gather(sc.forward(sc.local(s)))
            pp, pl, dX = sc.psi(sc.local(s))
            res["psi"] = np.array([pp, pl])
            res["deltaX"] = gather(dX)
            res["kinetic"] = np.array([sc.kinetic_term(sc.local(p0))])
            sf, pf = sc.leapfrog(sc.local(s), sc.local(p0), 2, 1e-3)
            res["leap_s"], res["leap_p"] = gather(sf), gather(pf)
            res["device_draw"] = gather(sc.draw_momenta_device(5, 3))
        res["gradient"] = gather(sc.gradient_psi(sc.local(s)))
        sc.close()
        if info.rank == 0:
            with bc.Chain(bc.Params(device=info.local_rank, **kw)) as ch:
                ch.set_static(Power=P, nobs=nobs, noise=noise, window=window)
                ref = {}
                m = ch.hamiltonian_mass(s)[0]
                if res_mass is not None:
                    ref["mass_f"] = m
                if calc_h == 0:
                    km, pw = ch.measure_spectrum(s, 40)
                    ref["spectrum"] = np.concatenate([km, pw])
                    ref["fft_roundtrip"] = s
                    ref["convolve"] = ch.convolve_inv_corr(s, P)
                    ref["forward"] = ch.forward(s)
                    pp, pl, dX = ch.psi(s)
                    ref["psi"] = np.array([pp, pl])
                    ref["deltaX"] = dX
                    ref["kinetic"] = np.array([ch.kinetic_term(p0)])
                    sf, pf = ch.leapfrog(s, p0, 2, 1e-3)
                    ref["leap_s"], ref["leap_p"] = sf, pf
                    ref["device_draw"] = ch.draw_momenta_device(5, 3)
                ref["gradient"] = ch.gradient_psi(s)
            for k in ref:
                e = rel(res[k], ref[k])
                worst[f"{k}[calc_h={calc_h}]"] = e
                ok &= e < 1e-11
    if info.rank == 0:
        for k, e in worst.items():
            print(f"  {k:28s} rel L2 {e:.2e}")
        print(f"slab_check grid {N} on {info.world} ranks:", "OK" if ok else "MISMATCH", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    multi.finalize()
    return int(flag.item() != 0)


if __name__ == "__main__":
    sys.exit(main())

"""The slab-decomposed chain on ONE GPU: its ranks as host threads of this process (bgpu_slab_create_local), every
collective a host rendezvous + device copies.  NCCL refuses two ranks on one device, so this is how the slab code
path -- packed layouts, fused transposes into the peers' receive buffers, stencil / density / residual halos,
all-reduced scalars, the slab leapfrog with its shared run-away verdict -- runs on the single-GPU test box.  Every
quantity is compared with the single-GPU chain, which the parity tests pin to the reference."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _problem(N, amp, seed=3):
    from barcode_b200 import inputs
    L = inputs.box_length(N)
    rng = np.random.default_rng(seed)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    nobs = np.maximum(0.0, 1.0 + 0.3 * rng.standard_normal(n)).reshape(N, N, N)
    w = rng.standard_normal((N, N, N))
    Ph = np.maximum(P[:, :, :N // 2 + 1], 0)
    s = amp * np.fft.irfftn(np.fft.rfftn(w) * np.sqrt(Ph * n / L ** 3), s=(N, N, N), axes=(0, 1, 2))
    invPh = np.where(Ph > 0, 1.0 / np.where(Ph > 0, Ph, 1.0), 0.0)
    p0 = np.fft.irfftn(np.fft.rfftn(rng.standard_normal((N, N, N))) * np.sqrt(invPh), s=(N, N, N), axes=(0, 1, 2))
    drift = np.fft.irfftn(np.fft.rfftn(p0) * (L ** 3 / n) * Ph, s=(N, N, N), axes=(0, 1, 2))
    p0 *= 0.1 * np.abs(s).max() / (np.abs(drift).max() * 2e-3)
    return L, P, nobs, s, p0


@pytest.mark.parametrize("world,p2p,generic,sfmodel,mk,like,mass_type,amp", [
    (2, "1", "0", 1, 1, 1, 1, 0.5),    # fused transpose over the peers' receive buffers (TMA stores)
    (4, "1", "0", 1, 1, 1, 1, 0.5),
    (2, "0", "0", 1, 1, 1, 1, 0.5),    # packed send buffer + all-to-all
    (2, "0", "1", 1, 1, 1, 1, 0.5),    # the cp.async slab pass that 1024^3 runs on
    (2, "1", "0", 2, 1, 1, 1, 0.5),    # 2LPT/ALPT forward model: 4-plane stencil halo, cell-boundary plane
    (2, "1", "0", 1, 2, 0, 2, 0.5),    # TSC + Poisson: finite-difference product (2-plane halo), force-spectrum mass
    (2, "1", "0", 1, 1, 2, 3, 0.1),    # log-normal, mean-force mass
])
def test_local_slab_chain_matches_single_gpu_chain(world, p2p, generic, sfmodel, mk, like, mass_type, amp, monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from barcode_b200 import chain as bc, slab
    monkeypatch.setenv("BGPU_SLAB_P2P", p2p)
    monkeypatch.setenv("BGPU_FFT_SLAB_GENERIC", generic)
    N = 128
    L, P, nobs, s, p0 = _problem(N, amp)
    one = np.ones((N, N, N))
    for calc_h in (0, 1, 4):
        if calc_h == 4 and sfmodel != 1:
            continue   # single-GPU only (bgpu_slab_create says so)
        kw = dict(N1=N, L1=L, masskernel=mk, likelihood=like, rsd_model=(sfmodel == 1), calc_h=calc_h,
                  mass_type=mass_type, sfmodel=sfmodel, N_bin=40)

        def rank_work(r, group, kw=kw, calc_h=calc_h):
            sc = slab.SlabChain.create_local(bc.Params(**kw), r, group)
            try:
                sc.set_static(Power=sc.local(P), nobs=sc.local(nobs), noise=sc.local(one), window=sc.local(one))
                res = {"mass_f": sc.hamiltonian_mass(sc.local(s))[0]}
                if calc_h == 0:
                    km, pw = sc.measure_spectrum(sc.local(s), 40)
                    res["spectrum"] = np.concatenate([km, pw])
                    res["fft_roundtrip"] = sc.fft_c2r(sc.fft_r2c(sc.local(s)))
                    res["convolve"] = sc.convolve_inv_corr(sc.local(s), sc.local(P))
                    res["forward"] = sc.forward(sc.local(s))
                    pp, pl, dX = sc.psi(sc.local(s))
                    res["psi"] = np.array([pp, pl])
                    res["deltaX"] = dX
                    res["kinetic"] = np.array([sc.kinetic_term(sc.local(p0))])
                    sf, pf = sc.leapfrog(sc.local(s), sc.local(p0), 2, 1e-3)
                    res["leap_s"], res["leap_p"] = sf, pf
                    res["device_draw"] = sc.draw_momenta_device(5, 3)
                res["gradient"] = sc.gradient_psi(sc.local(s))
                return res
            finally:
                sc.close()

        parts = slab.run_local_ranks(world, rank_work)
        scalars = ("spectrum", "psi", "kinetic")
        got = {k: (parts[0][k] if k in scalars else np.concatenate([p[k].reshape(-1, N, N) for p in parts], 0))
               for k in parts[0]}
        for k in scalars:   # every rank holds the same all-reduced numbers
            if k in parts[0]:
                assert all(np.array_equal(parts[0][k], p[k]) for p in parts[1:]), k
        with bc.Chain(bc.Params(**kw)) as ch:
            ch.set_static(Power=P, nobs=nobs, noise=one, window=one)
            ref = {"mass_f": ch.hamiltonian_mass(s)[0]}
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
        for k, want in ref.items():
            assert rel_l2(got[k], want) < 1e-11, (k, calc_h)
        if calc_h == 0 and generic == "0" and mass_type == 1:
            # the device draw does not depend on the decomposition at all (same stream, same colouring arithmetic, and
            # -- on the TMA-staged passes both chains run -- the same transform arithmetic); with a mass that is itself
            # all-reduced (types 2 / 3: the binned likelihood-force spectrum) the agreement is to rounding, checked above
            assert np.array_equal(got["device_draw"].ravel(), ref["device_draw"].ravel())


@pytest.mark.parametrize("world,rsd", [(2, False), (4, True)])
def test_local_slab_sph_matches_single_gpu_chain(world, rsd):
    """The reference's shipped default on slabs: SPH spline mass assignment (the density halo widened by the spline's
    reach) and its exact adjoint calc_h = 2 (the gather reads the residual from the halo-extended tile)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from barcode_b200 import chain as bc, slab
    N = 128
    L, P, nobs, s, p0 = _problem(N, 0.5)
    one = np.ones((N, N, N))
    kw = dict(N1=N, L1=L, masskernel=3, likelihood=1, rsd_model=rsd, calc_h=2, mass_type=1, sfmodel=1)

    def rank_work(r, group):
        sc = slab.SlabChain.create_local(bc.Params(**kw), r, group)
        try:
            sc.set_static(Power=sc.local(P), nobs=sc.local(nobs), noise=sc.local(one), window=sc.local(one))
            pp, pl, dX = sc.psi(sc.local(s))
            return {"psi": np.array([pp, pl]), "deltaX": dX, "gradient": sc.gradient_psi(sc.local(s))}
        finally:
            sc.close()

    parts = slab.run_local_ranks(world, rank_work)
    got = {k: (parts[0][k] if k == "psi" else np.concatenate([p[k].reshape(-1, N, N) for p in parts], 0))
           for k in parts[0]}
    with bc.Chain(bc.Params(**kw)) as ch:
        ch.set_static(Power=P, nobs=nobs, noise=one, window=one)
        pp, pl, dX = ch.psi(s)
        ref = {"psi": np.array([pp, pl]), "deltaX": dX, "gradient": ch.gradient_psi(s)}
    for k, want in ref.items():
        assert rel_l2(got[k], want) < 1e-11, k


def test_local_slab_host_stream_momenta_match_single_gpu_chain():
    """bgpu_color_momenta on slabs (the reference's seed-compatible draw, HMC_momenta.cc:42-92): every rank colours its
    own k-space rows from the full white-noise grid; the assembled momenta are the single-GPU chain's."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from barcode_b200 import chain as bc, slab
    N, world = 128, 2
    L, P, nobs, s, p0 = _problem(N, 0.5)
    one = np.ones((N, N, N))
    rng = np.random.default_rng(11)
    white = rng.standard_normal((N, N, N)) + 1j * rng.standard_normal((N, N, N))
    kw = dict(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=0, mass_type=1, sfmodel=1)

    def rank_work(r, group):
        sc = slab.SlabChain.create_local(bc.Params(**kw), r, group)
        try:
            sc.set_static(Power=sc.local(P), nobs=sc.local(nobs), noise=sc.local(one), window=sc.local(one))
            sc.hamiltonian_mass(sc.local(s))
            return sc.color_momenta(white)
        finally:
            sc.close()

    parts = slab.run_local_ranks(world, rank_work)
    got = np.concatenate([p.reshape(-1, N, N) for p in parts], 0)
    with bc.Chain(bc.Params(**kw)) as ch:
        ch.set_static(Power=P, nobs=nobs, noise=one, window=one)
        ch.hamiltonian_mass(s)
        want = ch.color_momenta(white)
    assert rel_l2(got, want) < 1e-12

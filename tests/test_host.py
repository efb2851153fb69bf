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


def test_philox_known_answers():
    """Philox4x32-10 of the oracle (the generator behind bgpu_draw_momenta_device) against the published
    known-answer vectors of Random123 (kat_vectors: zero, all-ones and pi-digit inputs)."""
    from oracle import barcode_oracle as bo
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = bo.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_device_normals_oracle_statistics():
    from oracle import barcode_oracle as bo
    x = bo.device_normals(seed=12345, draw=7, stream=0, first=0, n=1 << 18)
    assert abs(x.mean()) < 5 / np.sqrt(x.size) and abs(x.var() - 1) < 5 * np.sqrt(2 / x.size)
    assert abs((x ** 4).mean() - 3) < 0.1
    # any sub-range regenerates independently; draws and streams differ
    assert np.array_equal(bo.device_normals(12345, 7, 0, 1000, 64), x[1000:1064])
    assert not np.array_equal(bo.device_normals(12345, 8, 0, 0, 64), x[:64])
    assert not np.array_equal(bo.device_normals(12345, 7, 1, 0, 64), x[:64])


def test_shared_x_pass_algebra():
    """The pass sequence behind BGPU_SHARE_X=1 (api.cu forward_from_shat / backproject, fft_ops.h K_MULK*,
    K_COMP_UNIT), emulated pass by pass with numpy, against the three separate transforms it replaces: k_y and k_z
    are constant along an x pencil, so the y and z components of a displacement / gradient / back-projection triple
    can share one x pass (same Nyquist and k^2 zeroing rules, EqSolvers.cc:208-268, gradient.cpp:38-74,167-210)."""
    from barcode_b200 import inputs
    N, L, a = 16, 50.0, -0.7
    rng = np.random.default_rng(3)
    k1 = inputs.calc_ki(N, L)
    nzh = N // 2 + 1
    kx, ky, kz = k1[:, None, None], k1[None, :, None], k1[None, None, :nzh]
    k2 = kx ** 2 + ky ** 2 + kz ** 2
    idx = np.arange(N)
    nyq = (idx[:, None, None] == N // 2) | (idx[None, :, None] == N // 2) | (idx[None, None, :nzh] == N // 2)
    rot = lambda v: v.imag - 1j * v.real                        # (Im v, -Re v)
    inv_disp = np.where((k2 > 1e-14) & ~nyq, a / np.where(k2 > 0, k2, 1.0), 0.0)
    inv_lap = np.where((k2 > 0) & ~nyq, 1.0 / np.where(k2 > 0, k2, 1.0), 0.0)
    kc = (kx, ky, kz)

    # displacement triple: Psi_c = IFFT[a k_c / k^2 (Im s^, -Re s^)]
    shat = np.fft.rfftn(rng.standard_normal((N, N, N)))
    B = np.fft.ifft(inv_disp * rot(shat), axis=0)                # the shared x pass, comp = K_COMP_UNIT
    for c in (1, 2):
        direct = np.fft.irfftn(kc[c] * inv_disp * rot(shat), s=(N, N, N), axes=(0, 1, 2))
        shared = np.fft.irfft(np.fft.ifft(kc[c] * B, axis=1), n=N, axis=2)   # K_MULK on the y pass's load, z pass
        assert np.abs(shared - direct).max() < 1e-13 * np.abs(direct).max()

    # gradient triple (gradfft): d_c = IFFT[-k_c (Im d^, -Re d^)], zero on the Nyquist planes
    G = np.fft.ifft(np.where(nyq, 0.0, -1.0) * rot(shat), axis=0)
    for c in (1, 2):
        direct = np.fft.irfftn(np.where(nyq, 0.0, -kc[c]) * rot(shat) + 0 * k2, s=(N, N, N), axes=(0, 1, 2))
        shared = np.fft.irfft(np.fft.ifft(kc[c] * G, axis=1), n=N, axis=2)
        assert np.abs(shared - direct).max() < 1e-13 * np.abs(direct).max()

    # back-projection triple: acc = sum_c k_c / k^2 (Im V^_c, -Re V^_c)
    V = [rng.standard_normal((N, N, N)) for _ in range(3)]
    direct = sum(kc[c] * inv_lap * rot(np.fft.rfftn(V[c])) for c in range(3))
    zy = lambda v: np.fft.fft(np.fft.rfft(v, axis=2), axis=1)    # z pass, y pass
    U = ky * zy(V[1])                                            # K_MULK_SET on the y pass's store
    U = U + kz * zy(V[2])                                        # K_MULK_ADD (TMA reduce-add)
    shared = kx * inv_lap * rot(np.fft.rfftn(V[0])) + inv_lap * rot(np.fft.fft(U, axis=0))   # K_INVLAP_ADD, unit
    assert np.abs(shared - direct).max() < 1e-13 * np.abs(direct).max()


def test_e2e_chains_leg_never_raises_without_a_gpu():
    """bench.py's interleaved-chains e2e leg runs in a child process with a timeout: where it cannot run (no GPU
    here) the bench line gets an error string, not an exception -- and no CPU fallback number."""
    import argparse
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible; the failure path is exercised on the CPU box")
    sys.path.insert(0, ROOT)
    import bench
    out = bench.e2e_chains_leg(argparse.Namespace(grid=32, calc_h=0, steps=2, warmup=1, chains=2), timeout_s=120)
    assert set(out) == {"error"} and "no CUDA device" in out["error"]


def test_kspace_leapfrog_algebra():
    """The trajectory in k-space (api.cu leapfrog_kspace, kernels.cu kspace_drift_kernel / KspaceKickF), emulated with
    numpy on the half grid, against the oracle's Hamiltonian_EoM (HMC.cc:251-369): with a Fourier-space mass the drift is
    s^ += eps (V/N)/M p^ and the kick p^ += a FFT[gradpsi], the half kicks between steps merge, s and p are transformed
    once at each end; momenta[0] for the run-away test is (1/N) sum_k w_k Re p^_k with w = 1 on the planes k_z = 0 and
    N/2 (whose mirrors are stored) and 2 elsewhere."""
    from oracle import barcode_oracle as bo
    N, L = 16, 50.0
    kw = dict(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=0, mass_type=1, sfmodel=1)
    p = bo.Params(**kw)
    rng = np.random.default_rng(8)
    kk = np.fft.fftfreq(N, d=L / N) * 2 * np.pi
    k2 = kk[:, None, None] ** 2 + kk[None, :, None] ** 2 + kk[None, None, :] ** 2
    power = np.where(k2 > 0, 2.0e3 / (1.0 + (np.sqrt(k2) / 0.2) ** 2), 0.0)
    s0 = 0.3 * rng.standard_normal((N, N, N))
    p0 = rng.standard_normal((N, N, N))
    nobs = np.maximum(0.0, 1.0 + 0.3 * rng.standard_normal((N, N, N)))
    ones = np.ones((N, N, N))
    mass_f = np.where(power > 0, 1.0 / np.where(power > 0, power, 1.0), 0.0)   # mass_type 1 with factor 1
    neps, eps = 3, 2e-3
    s_ref, p_ref = bo.leapfrog(p, s0, p0, neps, eps, power, nobs, ones, ones, mass_f, None)

    nzh = N // 2 + 1
    normFS = L ** 3 / N ** 3
    mf = mass_f[:, :, :nzh]
    inv_mass = np.where(mf > 0, normFS / np.where(mf > 0, mf, 1.0), 0.0)    # launch_inverse_spectrum
    shat, phat = np.fft.rfftn(s0), np.fft.rfftn(p0)
    w = np.full(nzh, 2.0)
    w[0] = w[-1] = 1.0

    def kick(a):
        g = bo.gradient_psi(p, np.fft.irfftn(shat, s=(N, N, N), axes=(0, 1, 2)), power, nobs, ones, ones)
        return phat + a * np.fft.rfftn(g)          # = p^ + a ((V/N)/P s^ + norm h^): gradpsi's last sum, untransformed

    phat = kick(-0.5 * eps)
    for j in range(neps):
        shat = shat + eps * inv_mass * phat
        phat = kick(-0.5 * eps if j + 1 == neps else -eps)
    s_k = np.fft.irfftn(shat, s=(N, N, N), axes=(0, 1, 2))
    p_k = np.fft.irfftn(phat, s=(N, N, N), axes=(0, 1, 2))
    assert np.abs(s_k - s_ref).max() < 1e-12 * np.abs(s_ref).max()
    assert np.abs(p_k - p_ref).max() < 1e-12 * np.abs(p_ref).max()
    p0_from_modes = float(np.sum(w * phat.real) / N ** 3)
    assert abs(p0_from_modes - p_ref.flat[0]) < 1e-12 * np.abs(p_ref).max()

"""The slab decomposition's host logic on CPU, world size 2 and 4 over gloo: the distributed FFT
(pack -> all-to-all -> transposed layout) and the halo-exchanged mass assignment, restated in numpy
around barcode_b200/slab.py's partition arithmetic (oracle/slab_oracle.py), reproduce the
single-process oracle."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from barcode_b200 import slab
        from oracle import barcode_oracle as bo, slab_oracle as so
        N = 16
        L = 50.0
        rng = np.random.default_rng(5)
        a = rng.standard_normal((N, N, N))
        x0, Ns = slab.slab_range(N, rank, world)
        # FFT: forward lands in the transposed layout, inverse returns to the x slab
        k = so.slab_rfftn(a[x0:x0 + Ns], N, rank, world, dist)
        ref = np.fft.rfftn(a)
        e_fwd = np.abs(k - ref[:, x0:x0 + Ns, :]).max() / np.abs(ref).max()
        back = so.slab_irfftn(k, N, rank, world, dist)
        e_inv = np.abs(back - a[x0:x0 + Ns]).max()
        # mass assignment with halo exchange, against the oracle's density of the full particle set
        p = bo.Params(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=0, mass_type=1)
        psi = p.d * rng.uniform(-1.5, 1.5, (3, N, N, N))        # crosses slab faces and the periodic box edge
        x, y, z = bo.positions(p, psi)
        rho_ref = bo.density(p, x, y, z)
        rho, H = so.slab_density(p, psi[:, x0:x0 + Ns], rank, world, dist)
        e_rho = np.abs(rho - rho_ref[x0:x0 + Ns]).max()
        # 2LPT/ALPT pieces that reach across slab faces: Hessian minors (4-plane halo), cell-boundary averaging
        phi = rng.standard_normal((N, N, N))
        e_m2v = np.abs(so.slab_calc_m2v(p, phi[x0:x0 + Ns], rank, world, dist) - bo.calc_m2v(p, phi)[x0:x0 + Ns]).max() \
            / np.abs(bo.calc_m2v(p, phi)).max()
        cb_ref = 0.5 * (np.roll(psi[0], (1, 1, 1), (0, 1, 2)) + psi[0])
        e_cb = np.abs(so.slab_cellbound(psi[0][x0:x0 + Ns], rank, world, dist) - cb_ref[x0:x0 + Ns]).max()
        assert e_m2v < 1e-13 and e_cb < 1e-15, (e_m2v, e_cb)
        # exact CIC adjoint with the residual's halo, finite difference along x with its 2-plane halo, binned spectrum
        r = rng.standard_normal((N, N, N))
        V = so.slab_gather_adjoint_cic(p, r[x0:x0 + Ns], psi[:, x0:x0 + Ns], rank, world, dist)
        Vref = bo.gather_adjoint(p, r, x, y, z)
        e_V = max(np.abs(V[c] - Vref[c][x0:x0 + Ns]).max() for c in range(3)) / max(np.abs(v).max() for v in Vref)
        e_fd = np.abs(so.slab_findif_x(p, phi[x0:x0 + Ns], rank, world, dist) - bo.gradfindif(p, phi, 1)[x0:x0 + Ns]).max() \
            / np.abs(bo.gradfindif(p, phi, 1)).max()
        km, pw = so.slab_measure_spectrum(p, k, 12, rank, world, dist)
        km0, pw0 = bo.measure_spectrum(p, a, 12)
        e_sp = max(np.abs(km - km0).max() / km0.max(), np.abs(pw - pw0).max() / pw0.max())
        assert e_V < 1e-13 and e_fd < 1e-14 and e_sp < 1e-13, (e_V, e_fd, e_sp)
        # shared x pass on slabs (inverse direction): one transpose serves the y and the z component of a triple
        kk = np.fft.fftfreq(N, d=L / N) * 2 * np.pi
        nzh = N // 2 + 1
        ky, kz = kk, kk[:nzh]
        got_y, got_z = so.slab_shared_inverse_pair(k, ky, kz, N, rank, world, dist)
        want_y = so.slab_irfftn(ky[None, x0:x0 + Ns, None] * k, N, rank, world, dist)   # the separate transforms
        want_z = so.slab_irfftn(kz[None, None, :] * k, N, rank, world, dist)
        e_sh = max(np.abs(got_y - want_y).max(), np.abs(got_z - want_z).max()) / np.abs(want_y).max()
        # momenta[0] of a k-space trajectory: every rank's share of (1/N^3) sum_k p^_k, all-reduced
        e_p0 = abs(so.slab_momenta0(k, N, world, dist) - a.flat[0])
        assert e_sh < 1e-13 and e_p0 < 1e-13, (e_sh, e_p0)
        q.put((rank, e_fwd, e_inv, e_rho, H, float(rho.sum())))
    except Exception as exc:  # surface the failure instead of letting the parent wait for its timeout
        q.put((rank, repr(exc)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_decomposition_matches_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert len(r) == 6, r
    for p in procs:
        assert p.exitcode == 0
    total = 0.0
    for rank, e_fwd, e_inv, e_rho, H, s in res:
        assert e_fwd < 1e-14 and e_inv < 1e-14, (rank, e_fwd, e_inv)
        assert e_rho < 1e-12, (rank, e_rho)
        assert 2 <= H <= 16 // world
        total += s
    assert abs(total - 16 ** 3) < 1e-8      # unit masses: every particle is counted once across ranks


def test_partition_arithmetic():
    sys.path.insert(0, ROOT)
    from barcode_b200 import slab
    assert slab.slab_range(512, 3, 8) == (192, 64)
    with pytest.raises(ValueError):
        slab.slab_range(100, 0, 8)
    # packed layout: element (x_l, y, z) of rank r goes to block y // Ns, where it is row (x_l, y % Ns)
    N, Ns = 16, 4
    idx = slab.packed_index(N, Ns, 2, 9, 5)
    nzh = N // 2 + 1
    assert idx == ((2 * Ns + 2) * Ns + 1) * nzh + 5
    # halo rule and the extended-tile plane map (periodic)
    assert slab.halo_planes(0.0, 1.0) == 2 and slab.halo_planes(2.5, 1.0) == 5
    assert list(slab.ext_plane(np.array([14, 15, 0, 5]), 0, 2, 16)) == [0, 1, 2, 7]

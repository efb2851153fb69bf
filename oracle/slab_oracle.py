"""TEST INFRASTRUCTURE -- numpy restatement of the slab-decomposed chain's data movement.

The product (barcode_b200/csrc: fft_plan.cu r2c_impl / c2r_impl, api.cu forward_from_shat,
kernels.cu deposit) distributes the reference's 3-D FFT (fftwrapper.cc:26-125) and mass
assignment (massFunctions.cc:49-364) over x slabs; the reference itself has no distributed mode,
so the check is self-consistency: the distributed algorithm, run with numpy per rank and
torch.distributed for the exchanges (gloo on CPU), must reproduce the single-process oracle
(oracle/barcode_oracle.py).  The partition arithmetic is imported from the product
(barcode_b200/slab.py) -- that is the piece under test.
"""
from __future__ import annotations

import numpy as np

from barcode_b200 import slab
from oracle import barcode_oracle as bo


def _all_to_all(blocks, rank, world, dist):
    """blocks[h] goes to rank h; returns the list of blocks received, indexed by source rank."""
    import torch
    mine = torch.from_numpy(np.ascontiguousarray(np.stack(blocks)).view(np.float64))
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)   # gloo has no all_to_all: gather everything, keep what is addressed to me
    cplx = blocks[0].dtype == np.complex128
    res = []
    for src in range(world):
        a = out[src].numpy()[rank]
        res.append(a.view(np.complex128) if cplx else a)
    return res


def slab_rfftn(local, N, rank, world, dist):
    """Real x-slab [Ns][N][N] -> transposed k-space slab [x][y_local][z <= N/2]."""
    x0, Ns = slab.slab_range(N, rank, world)
    nzh = N // 2 + 1
    zy = np.fft.fft(np.fft.rfft(local, axis=2), axis=1)                      # local z and y passes
    packed = np.empty(world * Ns * Ns * nzh, dtype=np.complex128)            # the y pass stores packed
    xl, y, z = np.meshgrid(np.arange(Ns), np.arange(N), np.arange(nzh), indexing="ij")
    packed[slab.packed_index(N, Ns, xl, y, z)] = zy
    blocks = list(packed.reshape(world, Ns, Ns, nzh))
    recv = _all_to_all(blocks, rank, world, dist)                            # block src = x planes of rank src
    tr = np.concatenate(recv, axis=0)                                        # [x][y_local][z]
    return np.fft.fft(tr, axis=0)                                            # local x pass


def slab_irfftn(kslab, N, rank, world, dist):
    """Transposed k-space slab -> real x-slab (1/N^3 included, fftwrapper.cc:43-45)."""
    x0, Ns = slab.slab_range(N, rank, world)
    nzh = N // 2 + 1
    tr = np.fft.ifft(kslab, axis=0)                                          # x pass on [x][y_local][z]
    blocks = [np.ascontiguousarray(tr[h * Ns:(h + 1) * Ns]) for h in range(world)]
    recv = _all_to_all(blocks, rank, world, dist)                            # from rank src: [x_l][y_l of src][z]
    zy = np.concatenate(recv, axis=1)                                        # [x_l][y][z], y = src*Ns + y_l
    return np.fft.irfft(np.fft.ifft(zy, axis=1), n=N, axis=2)


def slab_shared_inverse_pair(kslab, kfac_y, kfac_z, N, rank, world, dist):
    """The shared x pass on slabs, inverse direction (fft_plan.cu xpass_shared_inverse_impl / c2r_yz_shared_impl):
    ONE x pass and ONE all-to-all of `kslab` (a transposed k-space slab that already carries the x-pass functor with
    k_c := 1), then per component the real multiplier on the y pass's load (K_MULK: k_y varies along the y pencil, k_z
    along z) and the y and z passes.  kfac_y: [N], kfac_z: [N/2+1].  Returns the two real x-slabs
    IFFT[k_y * kslab], IFFT[k_z * kslab] (1/N^3 included)."""
    x0, Ns = slab.slab_range(N, rank, world)
    tr = np.fft.ifft(kslab, axis=0)                                          # the shared x pass
    blocks = [np.ascontiguousarray(tr[h * Ns:(h + 1) * Ns]) for h in range(world)]
    recv = _all_to_all(blocks, rank, world, dist)                            # stays in the receive buffer ...
    zy = np.concatenate(recv, axis=1)                                        # [x_l][y][z]
    out = []
    for mult in (kfac_y[None, :, None], kfac_z[None, None, :]):             # ... and is read once per component
        out.append(np.fft.irfft(np.fft.ifft(mult * zy, axis=1), n=N, axis=2))
    return out


def slab_momenta0(phat_slab, N, world, dist):
    """momenta[0] of a k-space trajectory on slabs (api.cu leapfrog_kspace, kernels.cu KspaceKickF): every rank sums
    w_k Re p^_k / N^3 over ITS transposed slab (w = 1 on the planes k_z = 0 and N/2, 2 elsewhere), one all-reduce."""
    import torch
    nzh = N // 2 + 1
    w = np.full(nzh, 2.0)
    w[0] = w[-1] = 1.0
    t = torch.tensor([float(np.sum(w * phat_slab.real)) / N ** 3], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


def slab_density(p: bo.Params, psi_local, rank, world, dist):
    """Mass assignment of this rank's particles into a halo-extended tile, halos added into the
    neighbours; returns (rho of the owned planes [Ns][N][N], halo width H)."""
    import torch
    N, d, L = p.N1, p.d, p.L1
    x0, Ns = slab.slab_range(N, rank, world)
    # positions of my Lagrangian planes (bo.positions on the slab: global lattice coordinate of plane i)
    g = d * np.arange(N, dtype=np.float64) + 0.5 * d
    x = bo.pacman(g[x0:x0 + Ns, None, None] + psi_local[0], L)
    y = bo.pacman(g[None, :, None] + psi_local[1], L)
    z = bo.pacman(g[None, None, :] + psi_local[2], L)
    if p.rsd_model:
        vez = bo.c_pecvel(p.ascale, p.OM, p.OL) * psi_local[2]
        OC = 1.0 - p.OM - p.OL
        Hub = 100.0 * np.sqrt(p.OM / p.ascale / p.ascale / p.ascale + p.OL + OC / p.ascale / p.ascale)
        z = bo.pacman(z + vez * (1.0 / Hub / p.ascale), L)
    # halo width from the largest x displacement on any rank
    m = torch.tensor([float(np.abs(psi_local[0]).max())], dtype=torch.float64)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    H = slab.halo_planes(float(m.item()), d)
    assert H <= Ns, "halo wider than a slab"
    ext = np.zeros((Ns + 2 * H) * N * N)
    xs, ys, zs = x.ravel(), np.broadcast_to(y, x.shape).ravel(), np.broadcast_to(z, x.shape).ravel()
    assert p.masskernel == 1, "the slab oracle restates the CIC deposit"
    ok = bo._in_domain(p, xs, ys, zs, False)
    i0, i1, tx, dx = bo.cic_cells_weights(p, xs[ok])
    j0, j1, ty, dy = bo.cic_cells_weights(p, ys[ok])
    k0, k1, tz, dz = bo.cic_cells_weights(p, zs[ok])
    for ii, wx in ((i0, tx), (i1, dx)):
        li = slab.ext_plane(ii, x0, H, N)
        assert (li < Ns + 2 * H).all(), "a particle left the halo"
        for jj, wy in ((j0, ty), (j1, dy)):
            for kk, wz in ((k0, tz), (k1, dz)):
                np.add.at(ext, kk + N * (jj + N * li), (1.0 * wx) * wy * wz)
    ext = ext.reshape(Ns + 2 * H, N, N)
    # lower halo -> last planes of rank-1, upper halo -> first planes of rank+1
    halos = torch.from_numpy(np.stack([ext[:H], ext[Ns + H:]]))
    allh = [torch.empty_like(halos) for _ in range(world)]
    dist.all_gather(allh, halos)
    lo, hi = (rank - 1) % world, (rank + 1) % world
    own = ext[H:H + Ns].copy()
    own[Ns - H:] += allh[hi][0].numpy()   # rank+1's lower halo are my last planes
    own[:H] += allh[lo][1].numpy()        # rank-1's upper halo are my first planes
    return own, H


def _neighbour_planes(local, nplanes, rank, world, dist):
    """(planes below my slab, planes above it): the last `nplanes` of rank-1 and the first of rank+1."""
    import torch
    edges = torch.from_numpy(np.ascontiguousarray(np.stack([local[:nplanes], local[-nplanes:]])))
    alle = [torch.empty_like(edges) for _ in range(world)]
    dist.all_gather(alle, edges)
    lo, hi = (rank - 1) % world, (rank + 1) % world
    return alle[lo][1].numpy(), alle[hi][0].numpy()


def slab_calc_m2v(p: bo.Params, phi_local, rank, world, dist):
    """calc_m2v (EqSolvers.cc:373-422, the 4th-order stencil applied twice) on an x slab: four halo planes
    from each x neighbour (api.cu forward_from_shat / kernels.cu lpt2_source_kernel), periodic in y and z."""
    N = p.N1
    XH = 4
    below, above = _neighbour_planes(phi_local, XH, rank, world, dist)
    ext = np.concatenate([below, phi_local, above], axis=0)          # planes [-4, Ns + 4)
    fac = N / (2.0 * p.L1)

    def fd(a, ax):  # gradient.cpp:81-153; along x no wrap (halo planes are there), the result loses 2 planes a side
        if ax == 0:
            c = a[2:-2]
            return -fac * ((4.0 / 3) * (a[1:-3] - a[3:-1]) - (1.0 / 6) * (a[:-4] - a[4:]))
        return -fac * ((4.0 / 3) * (np.roll(a, 1, ax) - np.roll(a, -1, ax))
                       - (1.0 / 6) * (np.roll(a, 2, ax) - np.roll(a, -2, ax)))

    def second(a, b):  # d_b d_a phi on the owned planes
        g = fd(ext, a)                      # a == 0: planes [-2, Ns+2); else all [-4, Ns+4)
        h = fd(g, b)
        lost = (2 if a == 0 else 0) + (2 if b == 0 else 0)
        return h[XH - lost:h.shape[0] - (XH - lost)]

    Lxx, Lxy, Lxz = second(0, 0), second(0, 1), second(0, 2)
    Lyy, Lyz, Lzz = second(1, 1), second(1, 2), second(2, 2)
    return Lxx * Lyy - Lxy * Lxy + Lxx * Lzz - Lxz * Lxz + Lyy * Lzz - Lyz * Lyz


def slab_cellbound(psi_local, rank, world, dist):
    """cellboundcomp (massFunctions.cc:588-660) on an x slab: 0.5 (Psi(i-1, j-1, k-1) + Psi(i, j, k)); plane
    x0 - 1 is the lower neighbour's last plane (kernels.cu scatter_kernel, GridGeom::cb_lo)."""
    below, _ = _neighbour_planes(psi_local, 1, rank, world, dist)
    ext = np.concatenate([below, psi_local], axis=0)
    shifted = np.roll(ext[:-1], (1, 1), (1, 2))                      # (i-1) by the halo, (j-1, k-1) periodic
    return 0.5 * (shifted + psi_local)


def slab_gather_adjoint_cic(p: bo.Params, r_local, psi_local, rank, world, dist):
    """Exact CIC adjoint on an x slab (api.cu gradient_device, kernels.cu gather_adjoint_kernel): the residual of my
    planes is extended by H planes of each x neighbour (the halo width of the scatter), and my particles gather
    from that tile through slab.ext_plane.  Returns V (3 arrays over my Lagrangian planes)."""
    import torch
    N, d, L = p.N1, p.d, p.L1
    x0, Ns = slab.slab_range(N, rank, world)
    g = d * np.arange(N, dtype=np.float64) + 0.5 * d
    x = bo.pacman(g[x0:x0 + Ns, None, None] + psi_local[0], L)
    y = bo.pacman(g[None, :, None] + psi_local[1], L) + 0 * x
    z = bo.pacman(g[None, None, :] + psi_local[2], L) + 0 * x
    if p.rsd_model:
        vez = bo.c_pecvel(p.ascale, p.OM, p.OL) * psi_local[2]
        OC = 1.0 - p.OM - p.OL
        Hub = 100.0 * np.sqrt(p.OM / p.ascale / p.ascale / p.ascale + p.OL + OC / p.ascale / p.ascale)
        z = bo.pacman(z + vez * (1.0 / Hub / p.ascale), L)
    m = torch.tensor([float(np.abs(psi_local[0]).max())], dtype=torch.float64)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    H = slab.halo_planes(float(m.item()), d)
    below, above = _neighbour_planes(r_local, H, rank, world, dist)
    ext = np.concatenate([below, r_local, above], axis=0).reshape(-1)      # planes [x0 - H, x0 + Ns + H)
    xs, ys, zs = x.ravel(), y.ravel(), z.ravel()
    i0, i1, tx, dx = bo.cic_cells_weights(p, xs)
    j0, j1, ty, dy = bo.cic_cells_weights(p, ys)
    k0, k1, tz, dz = bo.cic_cells_weights(p, zs)
    V = [np.zeros_like(xs) for _ in range(3)]
    for ii, wx, gx in ((i0, tx, -1.0 / d), (i1, dx, 1.0 / d)):
        li = slab.ext_plane(ii, x0, H, N)
        assert (li < Ns + 2 * H).all(), "a particle left the halo"
        for jj, wy, gy in ((j0, ty, -1.0 / d), (j1, dy, 1.0 / d)):
            for kk, wz, gz in ((k0, tz, -1.0 / d), (k1, dz, 1.0 / d)):
                rc = ext[kk + N * (jj + N * li)]
                V[0] += rc * gx * wy * wz
                V[1] += rc * wx * gy * wz
                V[2] += rc * wx * wy * gz
    if p.rsd_model:
        V[2] = V[2] + bo.fgrow(p.ascale, p.OM, p.OL) * V[2]
    return [v.reshape(x.shape) for v in V]


def slab_findif_x(p: bo.Params, a_local, rank, world, dist):
    """gradfindif along x (gradient.cpp:81-153) on an x slab: two halo planes from each neighbour
    (kernels.cu findif_product_kernel with xo = 2)."""
    below, above = _neighbour_planes(a_local, 2, rank, world, dist)
    e = np.concatenate([below, a_local, above], axis=0)
    fac = p.N1 / (2.0 * p.L1)
    return -fac * ((4.0 / 3) * (e[1:-3] - e[3:-1]) - (1.0 / 6) * (e[:-4] - e[4:]))


def slab_measure_spectrum(p: bo.Params, kslab, n_bin, rank, world, dist):
    """measure_spectrum (field_statistics.cpp:20-90) from the transposed k-space slab [x][y_local][z <= N/2]: modes with
    0 < z < N/2 count twice (their mirror is not stored), bins are summed over the ranks, then normalised
    (kernels.cu spectrum_bin_kernel / spectrum_finish_kernel)."""
    import torch
    N = p.N1
    y0, Ns = slab.slab_range(N, rank, world)
    k = bo.calc_ki(N, p.L1)
    nzh = N // 2 + 1
    ktot = np.sqrt((k[:, None, None] ** 2 + k[None, y0:y0 + Ns, None] ** 2) + k[None, None, :nzh] ** 2)
    dk = np.sqrt(3.0 * k[N // 2] * k[N // 2]) / float(n_bin)
    b = (ktot / dk).astype(np.int64)
    w = np.where((np.arange(nzh) > 0) & (np.arange(nzh) < N // 2), 2.0, 1.0)[None, None, :] + 0 * ktot
    ok = b < n_bin
    acc = np.stack([np.bincount(b[ok], weights=(w * np.abs(kslab) ** 2)[ok], minlength=n_bin),
                    np.bincount(b[ok], weights=(w * ktot)[ok], minlength=n_bin),
                    np.bincount(b[ok], weights=w[ok], minlength=n_bin)])
    t = torch.from_numpy(acc)
    dist.all_reduce(t)
    power, kmode, nmode = t.numpy()
    have = nmode > 0
    kmode[have] /= nmode[have]
    power[have] = power[have] / nmode[have] * (p.vol / float(N) ** 3 / float(N) ** 3)
    return kmode, power

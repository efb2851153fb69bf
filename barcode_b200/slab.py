"""One HMC chain across G GPUs: x-slab decomposition (BASELINE.json configs[3], [4]).

Rank r owns planes i in [r*N/G, (r+1)*N/G) of every real-space array -- the contiguous sub-array
[N/G][N][N] of the reference's layout (disp_part.cc:60), hence of its output files -- and y rows
[r*N/G, ...) of every k-space array ("transposed" layout [x][y_local][z <= N/2]).  The device
side (barcode_b200/csrc: fft_plan.cu r2c_impl / c2r_impl, api.cu forward_from_shat) does

  * the 3-D FFT as local z and y passes, one all-to-all, local x pass (and the reverse); the y
    pass writes / reads the packed all-to-all buffer [peer][x_local][y_local][z] directly,
  * the mass assignment into a density tile with H halo planes each side, H from the largest
    x displacement on any rank, halos added into the two x neighbours,
  * all-reduced scalars (sum rho, -lnL, prior, kinetic energy).

This module holds the host-side pieces: the partition arithmetic (shared with the numpy slab
oracle in oracle/slab_oracle.py and tested on CPU under gloo), the NCCL bootstrap through
torch.distributed, and `SlabChain`, the slab view of `Chain`.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .chain import Chain, Params


# ----------------------------------------------------------------------------- partition arithmetic
def slab_range(N: int, rank: int, world: int):
    """(x0, Ns): first plane and number of planes of `rank` (N must divide evenly, as the FFT needs)."""
    if N % world:
        raise ValueError(f"grid size {N} is not a multiple of the number of ranks {world}")
    ns = N // world
    return rank * ns, ns


def packed_index(N: int, Ns: int, xl, y, z):
    """Offset (in complex elements) of element (x_local, y, z) of an x-slab [Ns][N][N/2+1] inside the
    packed all-to-all send buffer [peer][x_local][y_local][z]; peer = y // Ns receives block `peer`."""
    nzh = N // 2 + 1
    peer, yl = np.divmod(y, Ns)
    return ((peer * Ns + xl) * Ns + yl) * nzh + z


def halo_planes(max_abs_psi_x: float, d: float) -> int:
    """Halo width each side of the density tile: ceil(max |Psi_x| / d) + 2 (one plane for the upper
    CIC / TSC neighbour, one for the lower TSC neighbour and rounding) -- api.cu forward_from_shat."""
    return int(math.ceil(max_abs_psi_x / d)) + 2


def ext_plane(cell_plane, x0: int, H: int, N: int):
    """Plane of the halo-extended local density tile [(Ns + 2H)][N][N] that global plane `cell_plane`
    maps to (periodic); >= Ns + 2H means beyond the halo (kernels.cu deposit())."""
    return np.mod(np.asarray(cell_plane) - (x0 - H), N)


# ----------------------------------------------------------------------------- NCCL bootstrap
def nccl_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    _lib.check(_lib.load().bgpu_nccl_unique_id(buf))
    return bytes(buf)


def broadcast_unique_id(rank: int, world: int) -> bytes:
    """Rank 0 draws the NCCL unique id; torch.distributed (any backend) hands it to the others."""
    import torch.distributed as dist
    obj = [nccl_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    return obj[0]


# ----------------------------------------------------------------------------- the chain
class SlabChain(Chain):
    """`Chain` whose arrays are this rank's slab [Ns][N][N]; every rank makes the same calls."""

    def __init__(self, params: Params, rank: int, world: int, unique_id: bytes | None, local_group=None):
        self.params = params
        self.L = _lib.load()
        self.N1 = int(params.N1)
        self.rank, self.world = int(rank), int(world)
        self.x0, self.Ns = slab_range(self.N1, rank, world)
        self.N = self.Ns * self.N1 * self.N1
        self.Nhalf = self.Ns * self.N1 * (self.N1 // 2 + 1)
        self._h = C.c_void_p()
        cp = params.to_c()
        idbuf = None
        if local_group is not None:
            _lib.check(self.L.bgpu_slab_create_local(C.byref(cp), self.rank, self.world, local_group.ptr, C.byref(self._h)))
            return
        if world > 1:
            if unique_id is None or len(unique_id) != 128:
                raise ValueError("a 128-byte NCCL unique id is required")
            idbuf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _lib.check(self.L.bgpu_slab_create(C.byref(cp), self.rank, self.world, idbuf, C.byref(self._h)))

    @classmethod
    def create(cls, params: Params, rank: int, world: int):
        """Collective constructor: draws a fresh NCCL unique id on rank 0 (an id bootstraps exactly one
        communicator) and distributes it through the initialised torch.distributed group."""
        return cls(params, rank, world, broadcast_unique_id(rank, world) if world > 1 else None)

    @classmethod
    def create_local(cls, params: Params, rank: int, group: "LocalGroup"):
        """Rank `rank` of a chain whose ranks all live in this process on one device (bgpu_slab_create_local): call
        from the host thread that will drive this rank; collective across the group's threads."""
        return cls(params, rank, group.world, None, local_group=group)

    def fused_transpose(self) -> bool:
        """True where the distributed FFT's transpose rides inside the strided pass (TMA stores into the peers'
        receive buffers): the TMA-staged sizes, unless BGPU_SLAB_P2P=0 / BGPU_FFT_SLAB_GENERIC=1."""
        import os
        return (self.world > 1 and self.N1 in (128, 256, 512) and os.environ.get("BGPU_SLAB_P2P", "1") != "0"
                and os.environ.get("BGPU_FFT_SLAB_GENERIC", "0") != "1")

    @property
    def shape(self):
        return (self.Ns, self.N1, self.N1)

    @property
    def kshape(self):
        """k-space arrays live transposed: all x, this rank's y rows, z <= N/2."""
        return (self.N1, self.Ns, self.N1 // 2 + 1)

    def local(self, full):
        """This rank's slab of a full (N, N, N) array."""
        a = np.asarray(full).reshape(self.N1, self.N1, self.N1)
        return np.ascontiguousarray(a[self.x0:self.x0 + self.Ns])


class LocalGroup:
    """The rendezvous object of an in-process slab chain (bgpu_local_group_*): one per chain, shared by its ranks."""

    def __init__(self, world: int):
        self.world = int(world)
        self.ptr = C.c_void_p()
        _lib.check(_lib.load().bgpu_local_group_create(self.world, C.byref(self.ptr)))

    def close(self):
        if self.ptr:
            _lib.load().bgpu_local_group_destroy(self.ptr)
            self.ptr = C.c_void_p()


def run_local_ranks(world: int, fn):
    """Run fn(rank, group) on `world` host threads sharing one LocalGroup (ctypes releases the GIL inside the library, so
    the ranks' collective calls can meet); returns the list of results, re-raises the first exception."""
    import threading
    group = LocalGroup(world)
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = fn(r, group)
        except BaseException as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    group.close()
    for e in err:
        if e is not None:
            raise e
    return out

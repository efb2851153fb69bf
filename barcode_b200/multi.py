"""Multi-GPU plumbing for independent chains (one process and one chain per GPU).

The hot path shards across GPUs as independent HMC chains (BASELINE.json configs[2]; the
reference itself only supports separate processes per chain, cmake/Modules/Options.cmake:34):
there is no data-path collective.  `torch.distributed` is used only to agree on seeds, to
barrier around timed regions and to reduce timings / energies for reporting, so everything
here runs unchanged on the `gloo` backend on CPU (tests/test_multi_gloo.py) and on `nccl`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int

    @property
    def is_root(self) -> bool:
        return self.rank == 0


def rank_info() -> RankInfo:
    """RANK / WORLD_SIZE / LOCAL_RANK as torchrun exports them (single process if absent)."""
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                    int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str, info: RankInfo, device=None):
    if info.world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=info.rank, world_size=info.world, **kw)


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()


def chain_seed(base_seed: int, rank: int) -> int:
    """Distinct, reproducible RNG seed per chain (the reference seeds gsl_rng_mt19937 from
    input.par's `seed`, main.cc:143; chains in separate directories use separate seeds)."""
    return int(base_seed) + 17 * int(rank)


def barrier(info: RankInfo):
    if info.world > 1:
        dist.barrier()


def max_over_ranks(value: float, info: RankInfo, device="cpu") -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if info.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, info: RankInfo, device="cpu") -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if info.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def bind_to_gpu_numa(local_rank: int):
    """Pin this process (and the pinned host buffers it allocates from now on: first touch) to the CPUs of the NUMA
    node its GPU hangs off.  One process per GPU each staging N^3-cell arrays through host memory contend for one
    node's memory controllers otherwise (round 1: end-to-end efficiency 0.44 at 8 GPUs with every rank on node 0).
    Returns a dict describing what was done; never raises (containers may hide sysfs)."""
    info = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(f"{base}/numa_node") as f:
            node = int(f.read().strip())
        with open(f"{base}/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        info.update({"pci": bdf, "numa_node": node, "cpus": len(cpus), "allowed": len(allowed)})
        if cpus and node >= 0:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)
    return info


def gather_to_root(values, info: RankInfo, device="cpu"):
    """Per-chain scalars (energies, acceptance flags) to rank 0: [world, len(values)] or None."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if info.world == 1:
        return t[None].cpu()
    out = [torch.empty_like(t) for _ in range(info.world)]
    dist.all_gather(out, t)
    return torch.stack(out).cpu() if info.is_root else None


def aggregate_throughput(local_ms: float, units_per_rank: int, info: RankInfo, device="cpu") -> float:
    """Whole-job rate: all ranks' units divided by the slowest rank's device time."""
    ms = max_over_ranks(local_ms, info, device)
    return info.world * units_per_rank / (ms * 1e-3)

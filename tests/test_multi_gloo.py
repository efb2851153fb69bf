"""world_size-2 `gloo` run of the chain-sharding plumbing (barcode_b200/multi.py) on CPU."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from barcode_b200 import multi
    info = multi.rank_info()
    multi.init("gloo", info)
    seed = multi.chain_seed(1, info.rank)
    multi.barrier(info)
    local_ms = 10.0 * (rank + 1)                       # rank 1 is the slow one
    rate = multi.aggregate_throughput(local_ms, 5, info)
    total = multi.sum_over_ranks(float(seed), info)
    table = multi.gather_to_root([rank, seed, local_ms], info)
    q.put((rank, seed, rate, total, None if table is None else table.tolist()))
    multi.finalize()


def test_two_rank_chain_sharding():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seeds = [r[1] for r in res]
    assert seeds == [1, 18] and len(set(seeds)) == world          # distinct chains
    for r in res:
        assert abs(r[2] - world * 5 / 20e-3) < 1e-9               # all units / slowest rank's time
        assert r[3] == sum(seeds)
    assert res[0][4] == [[0.0, 1.0, 10.0], [1.0, 18.0, 20.0]] and res[1][4] is None


def test_single_process_defaults():
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    from barcode_b200 import multi
    info = multi.rank_info()
    assert (info.rank, info.world, info.local_rank) == (0, 1, 0) and info.is_root
    assert multi.aggregate_throughput(4.0, 8, info) == 8 / 4e-3
    assert multi.gather_to_root([1.0, 2.0], info).tolist() == [[1.0, 2.0]]

"""Host-side multi-GPU logic on CPU: utterance sharding + the bench's max-over-ranks reduction under gloo
(world_size 2).  The data path has no collective (SURVEY.md §8 e); only the timing does."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nhans_b200.runtime import shard_by_load, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 255, 8192):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_load_balances_ragged_batches():
    rng = np.random.default_rng(0)
    lengths = rng.integers(16000, 160000, 64).tolist()
    shards = shard_by_load(lengths, 4)
    assert sorted(i for s in shards for i in s) == list(range(64))
    loads = [sum(lengths[i] for i in s) for s in shards]
    assert max(loads) / min(loads) < 1.1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(101, rank, world)
    mine = torch.tensor([float(hi - lo)])
    tot = mine.clone()
    dist.all_reduce(tot)                                   # every utterance is owned exactly once
    ms = torch.tensor([10.0 + rank])                       # bench.py: time = max over ranks
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        out.put((float(tot), float(ms)))
    dist.destroy_process_group()


def test_gloo_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    tot, ms = q.get(timeout=10)
    assert tot == 101.0 and ms == 11.0

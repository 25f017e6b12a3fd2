"""Host-side multi-GPU logic on CPU: utterance sharding + the bench's max-over-ranks reduction under gloo
(world_size 2).  The data path has no collective (SURVEY.md §8 e); only the timing does."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nhans_b200.runtime import ChunkQueue, MultiGpu, make_chunks, shard_by_load, shard_range


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
    ms = torch.tensor([10.0 + rank])                       # bench.py: time = max over ranks, every rank's time is reported
    allv = [torch.zeros_like(ms) for _ in range(world)]
    dist.all_gather(allv, ms)
    assert [float(v) for v in allv] == [10.0 + r for r in range(world)]
    ms = torch.tensor([max(float(v) for v in allv)])
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


def test_make_chunks_cover_everything_once():
    rng = np.random.default_rng(1)
    for n, c in ((0, 4), (1, 64), (10, 3), (257, 64), (8192, 64)):
        lengths = rng.integers(400, 160000, n).tolist()
        chunks = make_chunks(lengths, c)
        assert sorted(i for ch in chunks for i in ch) == list(range(n))
        assert all(1 <= len(ch) <= c for ch in chunks)
        if n:
            # longest utterances go out first
            assert max(lengths) in [lengths[i] for i in chunks[0]]
    q = ChunkQueue(3)
    assert [q.take() for _ in range(5)] == [0, 1, 2, None, None]


class _FakeEngine:
    """Stands in for Engine in the dynamic deal: `delay` seconds of 'GPU time' per chunk."""

    def __init__(self, device, delay):
        self.device, self.delay, self.in_flight = device, delay, 0

    def submit(self, mixes, a, b, **kw):
        assert self.in_flight < 2                              # the library keeps at most two batches in flight
        self.in_flight += 1
        return dict(mixes=mixes, t=__import__("time").perf_counter())

    def collect(self, ticket, newer_in_flight=False):
        import time
        time.sleep(self.delay)
        self.in_flight -= 1
        return {"out_offs": None, "i16": [np.asarray(m) * 2 for m in ticket["mixes"]]}

    def close(self):
        pass


def test_dynamic_deal_follows_gpu_speed_and_keeps_order():
    """A GPU that is 4x slower ends up with about a quarter of the chunks; results come back in input order."""
    mg = MultiGpu.__new__(MultiGpu)
    mg.engines = [_FakeEngine(0, 0.002), _FakeEngine(1, 0.008)]
    mg.last_stats = None
    rng = np.random.default_rng(2)
    clips = [rng.integers(-100, 100, int(n)).astype(np.int16) for n in rng.integers(400, 2000, 300)]
    out = mg.enhance(clips, None, clips, chunk_utts=5)
    assert all(np.array_equal(out["i16"][i], clips[i] * 2) for i in range(300))
    st = mg.last_stats["per_gpu"]
    assert st[0]["chunks"] + st[1]["chunks"] == 60 and sum(s["utterances"] for s in st) == 300
    assert st[0]["chunks"] > 2 * st[1]["chunks"]               # the fast engine pulled most of the queue


def test_pass_aware_chunks_fill_whole_passes():
    """Chunks are closed where their windows nearly fill a whole number of 2048-window passes."""
    from nhans_b200.runtime import frames_of
    lengths = [160000] * 8192                                   # BASELINE config 5: 998 windows per clip
    chunks = make_chunks(lengths, 32, pass_windows=2048)
    assert sorted(i for ch in chunks for i in ch) == list(range(8192))
    waste = []
    for ch in chunks[:-1]:
        w = sum(frames_of(lengths[i]) for i in ch)
        passes = -(-w // 2048)
        waste.append(1.0 - w / (passes * 2048.0))
        assert 16 <= len(ch) <= 48
    assert max(waste) < 0.05 and sum(waste) / len(waste) < 0.03
    assert len(chunks) >= 8 * 20                                # fine enough for the dynamic deal on 8 GPUs
    rng = np.random.default_rng(3)
    ragged = rng.integers(8000, 200000, 500).tolist()
    ch = make_chunks(ragged, 16, pass_windows=2048)
    assert sorted(i for c in ch for i in c) == list(range(500))

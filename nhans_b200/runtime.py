"""Utterance sharding across the GPUs of one box (SURVEY.md §8 e): every utterance is independent in eval
mode (batch-norm uses population statistics, N_HANS___Selective_Noise/blocks.py:104-108), so the batch is
cut into contiguous blocks, one per GPU, each GPU holds a full weight replica and there is NO collective
on the data path - only host scatter of int16 PCM and host gather of int16 PCM."""
from __future__ import annotations

import threading


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_load(lengths, world):
    """Ragged batches: longest-first round-robin deal (balances frames per GPU). -> list of index lists."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += lengths[i]
    return [sorted(s) for s in out]


class MultiGpu:
    """One Engine per visible GPU, each driven by its own host thread (ctypes releases the GIL)."""

    def __init__(self, devices, variant, weights, win_capacity=0, row_capacity=0):
        from .engine import Engine
        self.engines = []
        for d in devices:
            e = Engine(d, variant, win_capacity, row_capacity)
            e.load_weights(weights)
            self.engines.append(e)

    def enhance(self, mix_clips, ctx_a_clips, ctx_b_clips, **kw):
        world = len(self.engines)
        shards = shard_by_load([len(c) for c in mix_clips], world)
        results = [None] * world
        errors = []

        def work(r):
            idx = shards[r]
            if not idx:
                return
            try:
                results[r] = self.engines[r].enhance([mix_clips[i] for i in idx],
                                                     None if ctx_a_clips is None else [ctx_a_clips[i] for i in idx],
                                                     [ctx_b_clips[i] for i in idx], **kw)
            except Exception as ex:  # surfaced to the caller below
                errors.append(ex)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        out = {}
        for r, idx in enumerate(shards):
            if not idx:
                continue
            for key, vals in results[r].items():
                if key == "out_offs":
                    continue
                out.setdefault(key, [None] * len(mix_clips))
                for j, i in enumerate(idx):
                    out[key][i] = vals[j]
        return out

    def close(self):
        for e in self.engines:
            e.close()

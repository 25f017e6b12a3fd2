"""Utterance sharding across the GPUs of one box (SURVEY.md §8 e): every utterance is independent in eval
mode (batch-norm uses population statistics, N_HANS___Selective_Noise/blocks.py:104-108), so a batch is cut into
chunks of utterances, each GPU holds a full weight replica and there is NO collective on the data path - only
host scatter of int16 PCM and host gather of int16 PCM.

Chunks are dealt DYNAMICALLY: every GPU has a host thread that pulls the next chunk from a shared queue as soon as
it has room, so a GPU that runs slower under the power cap (round 1 measured 2.7 % between boxes) simply takes
fewer chunks instead of holding the whole job back.  Each thread keeps two chunks in flight on its engine (pinned
double-buffered staging + the library's copy streams): chunk i + 1 is packed and uploaded while chunk i computes,
chunk i is unpacked while chunk i + 1 computes."""
from __future__ import annotations

import threading
import time


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_load(lengths, world):
    """Static deal for ragged batches: longest-first onto the least loaded GPU. -> list of index lists."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += lengths[i]
    return [sorted(s) for s in out]


def frames_of(n_samples):
    """STFT frames = mask-network windows of a clip (frame 400, hop 160; SN/apply.py:368-369)."""
    return 1 + (n_samples - 400) // 160 if n_samples >= 400 else 0


def make_chunks(lengths, chunk_utts, pass_windows=0):
    """Work units of the dynamic deal: utterances sorted longest first (the long chunks go out first, the short ones
    fill the tail) in groups of about `chunk_utts`.  With `pass_windows` (the engine's windows per network pass) a
    chunk is closed where its window count nearly fills a whole number of passes - a chunk of 64 ten-second clips is
    31.2 passes, i.e. 2.5 % of its last pass idle; 41 clips are 19.98 - within [chunk_utts / 2, 3 chunk_utts / 2].
    -> list of index lists covering every utterance exactly once."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    chunk_utts = max(1, int(chunk_utts))
    if pass_windows <= 0:
        return [sorted(order[k:k + chunk_utts]) for k in range(0, len(order), chunk_utts)]
    chunks, cur, win = [], [], 0
    lo, hi = max(1, chunk_utts // 2), max(1, (3 * chunk_utts) // 2)
    for i in order:
        cur.append(i)
        win += frames_of(lengths[i])
        fill = (win % pass_windows) / float(pass_windows)
        if len(cur) >= hi or (len(cur) >= lo and (fill >= 0.95 or fill == 0.0)):
            chunks.append(sorted(cur))
            cur, win = [], 0
    if cur:
        chunks.append(sorted(cur))
    return chunks


class ChunkQueue:
    """Thread-safe dispenser of chunk indices."""

    def __init__(self, n):
        self.n, self.next, self.lock = n, 0, threading.Lock()

    def take(self):
        with self.lock:
            if self.next >= self.n:
                return None
            k = self.next
            self.next += 1
            return k


class MultiGpu:
    """One Engine per visible GPU, each driven by its own host thread (ctypes releases the GIL)."""

    def __init__(self, devices, variant, weights, win_capacity=0, row_capacity=0):
        from .engine import Engine
        self.engines = []
        for d in devices:
            e = Engine(d, variant, win_capacity, row_capacity)
            e.load_weights(weights)
            self.engines.append(e)
        self.pass_windows = win_capacity if win_capacity > 0 else 2048
        self.last_stats = None

    def enhance(self, mix_clips, ctx_a_clips, ctx_b_clips, chunk_utts=32, **kw):
        """Same result as Engine.enhance on one GPU (utterances are independent), gathered in input order.
        self.last_stats holds per-GPU chunk counts, audio seconds and busy time of the call."""
        world = len(self.engines)
        chunks = make_chunks([len(c) for c in mix_clips], chunk_utts, getattr(self, "pass_windows", 0))
        queue = ChunkQueue(len(chunks))
        done = [None] * len(chunks)
        errors = []
        stats = [dict(device=e.device, chunks=0, utterances=0, samples=0, busy_s=0.0) for e in self.engines]

        def work(r):
            eng, st = self.engines[r], stats[r]
            t0 = time.perf_counter()
            prev = None                                   # (chunk index, ticket) in flight
            try:
                while True:
                    k = queue.take()
                    cur = None
                    if k is not None:
                        idx = chunks[k]
                        cur = (k, eng.submit([mix_clips[i] for i in idx],
                                             None if ctx_a_clips is None else [ctx_a_clips[i] for i in idx],
                                             [ctx_b_clips[i] for i in idx], **kw))
                        st["chunks"] += 1
                        st["utterances"] += len(idx)
                        st["samples"] += sum(len(mix_clips[i]) for i in idx)
                    if prev is not None:
                        done[prev[0]] = eng.collect(prev[1], newer_in_flight=cur is not None)
                    prev = cur
                    if cur is None:
                        break
            except Exception as ex:  # surfaced to the caller below
                errors.append(ex)
            st["busy_s"] = time.perf_counter() - t0

        t_all = time.perf_counter()
        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        out = {}
        for k, idx in enumerate(chunks):
            for key, vals in done[k].items():
                if key == "out_offs":
                    continue
                out.setdefault(key, [None] * len(mix_clips))
                for j, i in enumerate(idx):
                    out[key][i] = vals[j]
        self.last_stats = dict(wall_s=time.perf_counter() - t_all, chunk_utts=chunk_utts, n_chunks=len(chunks), per_gpu=stats)
        return out

    def close(self):
        for e in self.engines:
            e.close()

"""Property tests (hypothesis) of the host-side logic above the C ABI: batching helpers, utterance sharding, the
frame / trim / output-length arithmetic of SN/apply.py:157-161 + tf.signal.stft / inverse_stft (oracle and the
library's host entry point), wav I/O of SN/apply.py:46-53."""
import ctypes

import numpy as np
from hypothesis import given, settings, strategies as st

from nhans_b200 import _lib
from nhans_b200.engine import pack, unpack
from nhans_b200.runtime import shard_by_load, shard_range
from nhans_b200.wavio import read_wav, write_wav
from oracle import nhans_oracle as O

lengths = st.lists(st.integers(0, 5000), min_size=1, max_size=40)


@settings(max_examples=60, deadline=None)
@given(lengths)
def test_pack_unpack_roundtrip(ls):
    rng = np.random.default_rng(sum(ls))
    clips = [rng.integers(-32768, 32767, n).astype(np.int16) for n in ls]
    data, offs = pack(clips)
    assert data.dtype == np.int16 and offs.dtype == np.int64 and offs[0] == 0 and offs[-1] == sum(ls)
    back = unpack(data, offs)
    assert all(np.array_equal(a, b) for a, b in zip(clips, back))


@settings(max_examples=80, deadline=None)
@given(st.integers(0, 10000), st.integers(1, 16))
def test_shard_range_is_a_balanced_partition(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(1, 200000), min_size=1, max_size=64), st.integers(1, 8))
def test_shard_by_load_covers_every_utterance_once(ls, world):
    shards = shard_by_load(ls, world)
    assert len(shards) == world and sorted(i for s in shards for i in s) == list(range(len(ls)))
    loads = [sum(ls[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(ls)                       # greedy longest-first bound


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 400000))
def test_frame_and_length_arithmetic(n):
    t = O.trim_len(n)
    if n < 400:
        assert t == n or t <= n                                       # no frames: nothing to process
    else:
        assert t <= n and n - t < 160 and (t - 400) % 160 == 0
        T = 1 + (t - 400) // 160
        assert T == 1 + (n - 400) // 160                              # trimming never drops a frame
        assert (T - 1) * 160 + 400 == t                               # inverse STFT returns the trimmed length


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(0, 100000), min_size=1, max_size=30))
def test_library_output_offsets_agree_with_the_oracle(ls):
    lib = _lib.load()
    offs = np.concatenate([[0], np.cumsum(ls)]).astype(np.int64)
    out = np.zeros(len(ls) + 1, np.int64)
    rc = lib.nhans_output_offsets(offs.ctypes.data_as(ctypes.c_void_p), len(ls), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    want = [O.trim_len(n) if n >= 400 else 0 for n in ls]
    assert list(np.diff(out)) == want


def test_wav_io_mono_stereo(tmp_path):
    rng = np.random.default_rng(0)
    mono = rng.integers(-30000, 30000, 4000).astype(np.int16)
    p = str(tmp_path / "m.wav")
    write_wav(p, mono)
    assert np.array_equal(read_wav(p), mono)
    stereo = np.stack([mono, mono[::-1]], axis=1)
    p2 = str(tmp_path / "s.wav")
    write_wav(p2, stereo)
    got = read_wav(p2)                                                # SN/apply.py:50-51: mean over channels
    want = stereo.astype(np.float64).mean(axis=1)
    assert got.dtype == np.float64 and np.array_equal(got, want)      # exact: never rounded back to int16
    silent = np.zeros((1000, 2), np.int16)                            # Silent.wav is stereo zeros
    p3 = str(tmp_path / "z.wav")
    write_wav(p3, silent)
    assert not read_wav(p3).any()

"""Variable inventory / checkpoint reader / seeded init (host logic, CPU)."""
import json
import os
import struct

import numpy as np
import pytest

from nhans_b200 import weights as W

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag,variant", [("sn", 0), ("ss", 1)])
def test_inventory_matches_reference_checkpoint_index(tag, variant):
    """Names, shapes and byte sizes of model()'s variables equal the reference's own checkpoint index
    (tests/golden/ckpt_index_*.json, extracted from trained_model/*.index)."""
    idx = json.load(open(os.path.join(GOLD, "ckpt_index_%s.json" % tag)))
    f32 = {k: v for k, v in idx.items() if v[0] == 1}
    inv = W.inventory(variant)
    assert set(inv) == set(f32)
    assert len(inv) == 571
    for k, shp in inv.items():
        assert list(shp) == f32[k][1], k
        assert int(np.prod(shp)) * 4 == f32[k][3], k
    assert W.n_params(variant) == 28999881
    assert sum(v[3] for v in f32.values()) == 115999524          # the git-LFS pointer's size
    offs = sorted((v[2], v[3]) for v in idx.values())
    assert all(offs[i][0] + offs[i][1] == offs[i + 1][0] for i in range(len(offs) - 1))   # contiguous


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference mount absent")
def test_index_reader_on_reference_files():
    e = W.read_bundle_index("/root/reference/N_HANS___Selective_Noise/trained_model/81448_0-1000000.index")
    idx = json.load(open(os.path.join(GOLD, "ckpt_index_sn.json")))
    assert {k: [v["dtype"], list(v["shape"]), v["offset"], v["size"]] for k, v in e.items()} == idx
    with pytest.raises(FileNotFoundError):                           # the data shard is an LFS pointer
        W.load_bundle("/root/reference/N_HANS___Selective_Noise/trained_model/81448_0-1000000")
    # like the reference (SN/apply.py:428-432), a checkpoint that cannot be restored is an error ...
    with pytest.raises(W.CheckpointMissing):
        W.load_or_init(0, "/root/reference/N_HANS___Selective_Noise/trained_model", allow_random=False)
    # ... unless random-init weights of the identical architecture are explicitly allowed
    w, src = W.load_or_init(0, "/root/reference/N_HANS___Selective_Noise/trained_model", allow_random=True)
    assert src == "random-init" and len(w) == 571


def _varint(n):
    out = b""
    while True:
        b = n & 0x7F
        n >>= 7
        out += bytes([b | (0x80 if n else 0)])
        if not n:
            return out


def _block(entries):
    body = b""
    for k, v in entries:                      # no prefix compression, single restart
        body += _varint(0) + _varint(len(k)) + _varint(len(v)) + k + v
    body += struct.pack("<I", 0) + struct.pack("<I", 1)
    return body


def _write_bundle(prefix, tensors):
    """Minimal TF tensor-bundle writer (one data block) to round-trip the reader."""
    data = b""
    entries = [(b"", b"\x08\x01")]
    for name in sorted(tensors):
        a = np.ascontiguousarray(tensors[name], "<f4")
        shape = b"".join(b"\x12" + _varint(len(d)) + d for d in (b"\x08" + _varint(s) for s in a.shape))
        proto = b"\x08\x01" + b"\x12" + _varint(len(shape)) + shape + b"\x20" + _varint(len(data)) + b"\x28" + _varint(a.nbytes)
        entries.append((name.encode(), proto))
        data += a.tobytes()
    blk = _block(entries)
    f = blk + b"\x00" + b"\x00" * 4
    meta_off = len(f)
    meta = _block([])
    f += meta + b"\x00" + b"\x00" * 4
    idx_off = len(f)
    idx = _block([(b"\xff", _varint(0) + _varint(len(blk)))])
    f += idx + b"\x00" + b"\x00" * 4
    footer = _varint(meta_off) + _varint(len(meta)) + _varint(idx_off) + _varint(len(idx))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    open(prefix + ".index", "wb").write(f + footer)
    open(prefix + ".data-00000-of-00001", "wb").write(data)


def test_bundle_roundtrip(tmp_path):
    w = W.seeded_init(0, 3)
    small = {k: w[k] for k in list(sorted(w))[:40]}
    _write_bundle(str(tmp_path / "ckpt-1"), small)
    back = W.load_bundle(str(tmp_path / "ckpt-1"))
    assert set(back) == set(small)
    for k in small:
        assert back[k].shape == small[k].shape and np.array_equal(back[k], small[k])
    assert W.find_checkpoint(str(tmp_path)) == str(tmp_path / "ckpt-1")


def test_seeded_init_is_deterministic_and_non_degenerate():
    a, b = W.seeded_init(0, 0), W.seeded_init(0, 0)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    c = W.seeded_init(0, 1)
    assert not np.array_equal(a["last_dense/w"], c["last_dense/w"])
    # the reference initialises these with stddev 0 (main.py:136,142,146,238) -> identity network
    for k in ("last_dense/w", "resblock1_1_conv1_noise_pos_emb/w", "resblock1_1_conv1_temb_dense3/w"):
        assert float(np.abs(a[k]).max()) > 0
    assert set(W.seeded_init(1, 0)) == set(W.inventory(1))

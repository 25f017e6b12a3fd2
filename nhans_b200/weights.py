"""Variable inventory, seeded initialisation and TF tensor-bundle reader for the N-HANS network.

The reference builds its variables in ``model()`` (N_HANS___Selective_Noise/main.py:98-242,
N_HANS___Source_Separation/main.py:99-253) from the primitives in blocks.py:23-108 and stores them
with a TF1 ``Saver`` as a tensor bundle (``trained_model/*.index`` + ``*.data-00000-of-00001``).
This module restates the *names and shapes* of those variables, provides the
non-degenerate seeded initialisation used when the trained blob is absent (the
reference's own initialisers give an identity network, main.py:136,142,146,238), and reads the
bundle format directly so no TensorFlow is needed.

numpy only; used by the product path (host side) and by the oracle.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

SELECTIVE_NOISE = 0
SEPARATOR = 1

# (kernel, stride, channels) of the eight conditioned residual blocks, main.py:218-229
MAIN_BLOCKS = [
    ("resblock1_1", 4, 1, 64), ("resblock1_2", 4, 1, 64),
    ("resblock2_1", 4, 2, 128), ("resblock2_2", 4, 1, 128),
    ("resblock3_1", 3, 2, 256), ("resblock3_2", 3, 1, 256),
    ("resblock4_1", 3, 2, 512), ("resblock4_2", 3, 1, 512),
]
# ((kh, kw), (sh, sw), channels) of the embedding tower, main.py:192-197
TOWER_BLOCKS = [
    ("noise_resblock1_1", (8, 4), (3, 2), 64), ("noise_resblock2_1", (8, 4), (3, 2), 128),
    ("noise_resblock3_1", (4, 4), (1, 1), 256), ("noise_resblock4_1", (4, 4), (1, 2), 512),
]
N_BINS = 201
WINDOW_FRAMES = 35
CONTEXT_FRAMES = 200
EMB_DIM = 512


def same_out(n, s):
    return -(-n // s)


def cond_names(variant):
    """Scope suffixes of the two conditioning projections: (ctx_a, ctx_b).

    SN: ctx_a = positive-noise context, ctx_b = negative-noise context (SN/main.py:142-148).
    SS: ctx_a = interference speaker ('_noise_emb'), ctx_b = target speaker ('_clean_emb')
    (SS/main.py:157-163)."""
    if variant == SELECTIVE_NOISE:
        return "_noise_pos_emb", "_noise_neg_emb"
    return "_noise_emb", "_clean_emb"


def _bn(shapes, scope, mask_shape):
    for v in ("beta", "gamma", "pop_mean", "pop_variance"):
        shapes[scope + "/" + v] = tuple(mask_shape)


def inventory(variant=SELECTIVE_NOISE):
    """OrderedDict name -> shape of every float32 variable ``model()`` creates."""
    sh = OrderedDict()
    # embedding tower (shared weights, scope 'embedding/')
    cin = 1
    for name, (kh, kw), _, c in TOWER_BLOCKS:
        p = "embedding/" + name
        sh[p + "_conv1/w"] = (kh, kw, cin, c)
        _bn(sh, p + "_conv1", (1, 1, 1, c))
        sh[p + "_conv2/w"] = (kh, kw, c, c)
        sh[p + "_conv2/b"] = (1, 1, 1, c)
        if cin != c:
            sh[p + "_transform/w"] = (1, 1, cin, c)
            sh[p + "_transform/b"] = (1, 1, 1, c)
        _bn(sh, p + "_addition", (1, 1, 1, c))
        cin = c
    # main network
    sa, sb = cond_names(variant)
    cin = 1
    for name, k, _, c in MAIN_BLOCKS:
        sh[name + "_conv1/w"] = (k, k, cin, c)
        _bn(sh, name + "_conv1", (1, 1, 1, c))
        sh[name + "_conv2/w"] = (k, k, c, c)
        sh[name + "_conv2/b"] = (1, 1, 1, c)
        for site in ("_conv1", "_conv2"):
            for s in (sa, sb):
                sh[name + site + s + "/w"] = (EMB_DIM, c)
                sh[name + site + s + "/b"] = (1, c)
            for tf_ in ("_temb", "_femb"):
                scope = name + site + tf_
                sh[scope + "_dense1/w"] = (1, 50)
                sh[scope + "_dense2/w"] = (50, 50)
                sh[scope + "_dense3/w"] = (50, c)
                _bn(sh, scope + scope + "_dense1", (1, 50))   # doubled scope, main.py:131,134
                _bn(sh, scope + scope + "_dense2", (1, 50))
        if cin != c:
            sh[name + "_transform/w"] = (1, 1, cin, c)
            sh[name + "_transform/b"] = (1, 1, 1, c)
        _bn(sh, name + "_addition", (1, 1, 1, c))
        cin = c
    sh["last_conv/w"] = (5, 1, 512, 512)
    _bn(sh, "last_conv", (1, 1, 1, 512))
    sh["last_dense/w"] = (26 * 512, N_BINS)
    sh["last_dense/b"] = (1, N_BINS)
    return sh


def n_params(variant=SELECTIVE_NOISE):
    return int(sum(int(np.prod(s)) for s in inventory(variant).values()))


def seeded_init(variant=SELECTIVE_NOISE, seed=0):
    """Non-degenerate random weights of the identical architecture.

    One ``default_rng(seed)`` stream consumed in sorted-name order, float32."""
    rng = np.random.default_rng(seed)
    shapes = inventory(variant)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        leaf = name.rsplit("/", 1)[1]
        scope = name.rsplit("/", 1)[0]
        if leaf == "w":
            if len(shp) == 4:
                fan_in = shp[0] * shp[1] * shp[2]
            else:
                fan_in = shp[0]
            gain = 2.0
            if scope.endswith("_transform") or scope.endswith("_emb"):
                gain = 1.0
            if scope.endswith("_conv2"):
                gain = 0.25            # keeps the residual stack from growing geometrically
            if scope.endswith("last_dense"):
                gain = 2.5e-3          # calibrated so that |out| (the log-magnitude residual) is O(1)
            std = np.sqrt(gain / fan_in)
            if scope.endswith("_dense1"):
                # input is range(n): keep pre-activations O(1) for the largest n
                std = 1.0 / 201.0 if "_femb" in scope else 1.0 / 35.0
            if scope.endswith("_dense3"):
                std = np.sqrt(1.0 / fan_in) * 0.5
            a = rng.normal(0.0, std, size=shp)
        elif leaf in ("b", "beta", "pop_mean"):
            a = rng.normal(0.0, 0.1, size=shp)
        elif leaf == "gamma":
            a = rng.uniform(0.8, 1.2, size=shp)
        elif leaf == "pop_variance":
            a = rng.uniform(0.5, 1.5, size=shp)
        else:  # pragma: no cover
            raise KeyError(name)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# TF tensor bundle (.index is a LevelDB-format table; .data-* holds raw little-endian tensors)
# ----------------------------------------------------------------------------------------------
_TABLE_MAGIC = 0xDB4775248B80FB57


def _varint(buf, pos):
    res = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        res |= (b & 0x7F) << shift
        if not b & 0x80:
            return res, pos
        shift += 7


def _read_block(data, off, size):
    blk = data[off:off + size]           # trailer (type + crc32c) follows; blocks are uncompressed
    if data[off + size] != 0:
        raise ValueError("compressed index blocks are not supported")
    n_restarts = struct.unpack("<I", blk[-4:])[0]
    end = len(blk) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    out = []
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        out.append((key, blk[pos:pos + vlen]))
        pos += vlen
    return out


def _parse_proto(buf):
    """Minimal protobuf wire parser -> {field: [values]} (varints as int, bytes as bytes)."""
    pos = 0
    out = {}
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack("<I", buf[pos:pos + 4])[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack("<Q", buf[pos:pos + 8])[0]
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def read_bundle_index(index_path):
    """Parse ``<prefix>.index`` -> OrderedDict name -> dict(dtype, shape, shard, offset, size).

    BundleEntryProto fields: 1 dtype (1 = float32, 3 = int32), 2 shape, 3 shard_id, 4 offset,
    5 size, 6 crc32c."""
    with open(index_path, "rb") as f:
        data = f.read()
    footer = data[-48:]
    if struct.unpack("<Q", footer[-8:])[0] != _TABLE_MAGIC:
        raise ValueError("%s is not a tensor-bundle index" % index_path)
    pos = 0
    _, pos = _varint(footer, pos)
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    entries = OrderedDict()
    for _, handle in _read_block(data, ioff, isize):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _read_block(data, boff, bsize):
            if key == b"":
                continue                      # BundleHeaderProto
            m = _parse_proto(val)
            dims = []
            if 2 in m:
                shp = _parse_proto(m[2][0])
                for d in shp.get(2, []):
                    dims.append(_parse_proto(d).get(1, [0])[0])
            entries[key.decode()] = dict(dtype=m.get(1, [0])[0], shape=tuple(dims),
                                         shard=m.get(3, [0])[0], offset=m.get(4, [0])[0],
                                         size=m.get(5, [0])[0])
    return entries


class CheckpointMissing(FileNotFoundError):
    """No usable trained checkpoint (directory absent, no .index, or the data shard is a git-LFS pointer)."""


def _is_lfs_pointer(path):
    try:
        if os.path.getsize(path) > 1024:
            return False
        with open(path, "rb") as f:
            return f.read(64).startswith(b"version https://git-lfs")
    except OSError:
        return False


def load_bundle(prefix):
    """Load every float32 tensor of a TF checkpoint ``prefix`` (no TensorFlow needed).

    Raises CheckpointMissing when the data shard is absent or a git-LFS pointer (the state of the reference
    mount) and ValueError when it exists but is shorter than the index says (a truncated
    download must not be mistaken for 'no checkpoint')."""
    entries = read_bundle_index(prefix + ".index")
    data_path = prefix + ".data-00000-of-00001"
    need = max(e["offset"] + e["size"] for e in entries.values())
    if not os.path.exists(data_path) or _is_lfs_pointer(data_path):
        raise CheckpointMissing("%s is absent or a git-LFS pointer (need %d bytes)" % (data_path, need))
    if os.path.getsize(data_path) < need:
        raise ValueError("%s is truncated: %d bytes, the index needs %d" % (data_path, os.path.getsize(data_path), need))
    out = {}
    with open(data_path, "rb") as f:
        for name, e in entries.items():
            if e["dtype"] != 1:
                continue
            f.seek(e["offset"])
            a = np.frombuffer(f.read(e["size"]), dtype="<f4").reshape(e["shape"])
            out[name] = np.ascontiguousarray(a)
    return out


def find_checkpoint(model_dir):
    """Return the checkpoint prefix inside ``model_dir`` (the reference hard-codes
    './trained_model/81448_0-1000000', SN/apply.py:430-432; SS: '81457_2-545000')."""
    if not os.path.isdir(model_dir):
        return None
    for f in sorted(os.listdir(model_dir)):
        if f.endswith(".index"):
            return os.path.join(model_dir, f[:-len(".index")])
    return None


def random_init_allowed():
    return os.environ.get("NHANS_ALLOW_RANDOM_INIT", "0") not in ("", "0")


def load_or_init(variant=SELECTIVE_NOISE, model_dir=None, seed=0, allow_random=None):
    """Weights from the trained checkpoint under ``model_dir``.

    The reference hard-fails when the checkpoint cannot be restored (SN/apply.py:428-432), and so does this:
    without a usable checkpoint CheckpointMissing is raised - unless seeded random-init weights of the identical
    architecture are explicitly allowed (``allow_random=True`` or NHANS_ALLOW_RANDOM_INIT=1; what the tests and
    the benchmark use, because the mounted checkpoints are git-LFS pointers).  A truncated data shard or a
    checkpoint with missing variables is always an error.

    Returns (weights, source) with source in {'checkpoint', 'random-init'}."""
    if allow_random is None:
        allow_random = random_init_allowed()
    prefix = find_checkpoint(model_dir) if model_dir else None
    why = "no checkpoint (*.index) under %r" % (model_dir,)
    if prefix:
        try:
            w = load_bundle(prefix)
            want = inventory(variant)
            missing = [n for n in want if n not in w or tuple(w[n].shape) != tuple(want[n])]
            if missing:
                raise KeyError("checkpoint lacks %d variables, e.g. %s" % (len(missing), missing[0]))
            return {n: w[n] for n in want}, "checkpoint"
        except CheckpointMissing as ex:
            why = str(ex)
    if not allow_random:
        raise CheckpointMissing(why + "; set NHANS_ALLOW_RANDOM_INIT=1 to run with seeded random-init weights of the "
                                "identical architecture instead")
    return seeded_init(variant, seed), "random-init"

"""The identity behind frame_conv_kernel / window_expand_kernel (nhans_b200/csrc/misc.cu), checked on the CPU with
torch's conv2d as the referee: the first convolution of the mask network evaluated on every 35-frame window
(what SN/apply.py:378 + SN/main.py:221 do) equals ONE convolution per frame of the zero-extended utterance, in four
variants that differ in which kernel rows fall outside the window (window rows 0, 33, 34)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import nhans_oracle as O

KH = KW = 4
PT, PB, PL, PR = 1, 2, 1, 2          # TF 'SAME' for k = 4, s = 1
WIN, HALF = 35, 17


def _windows(spec):
    return O.strided_crop(spec, WIN, 1)                                    # [T, 35, 201], zero padded


def test_per_frame_variants_reproduce_the_per_window_convolution():
    rng = np.random.default_rng(0)
    T, C = 47, 5
    spec = rng.standard_normal((T, 201)).astype(np.float32)
    w = rng.standard_normal((KH, KW, 1, C)).astype(np.float32)
    wt = torch.from_numpy(w).permute(3, 2, 0, 1).contiguous()              # [C, 1, kh, kw]
    # reference: every window on its own, TF SAME padding
    win = torch.from_numpy(_windows(spec))[:, None]                        # [T, 1, 35, 201]
    ref = F.conv2d(F.pad(win, (PL, PR, PT, PB)), wt)                       # [T, C, 35, 201]
    # per frame: zero-extended utterance (17 virtual frames on either side), one partial sum per kernel row
    ext = np.zeros((T + 2 * HALF + KH, 201), np.float32)                   # index k = vf + 17 (+ slack for the taps)
    ext[HALF:HALF + T] = spec
    x = torch.from_numpy(ext)[None, None]
    x = F.pad(x, (PL, PR, PT, 0))                                          # row k of the output reads ext rows k - 1 + i
    P = [F.conv2d(x, wt[:, :, i:i + 1, :])[0, :, i:i + T + 2 * HALF] for i in range(KH)]   # P[i][c, k, w]
    variants = [P[0] + P[1] + P[2] + P[3], P[1] + P[2] + P[3], P[0] + P[1] + P[2], P[0] + P[1]]
    for n in (0, 1, 16, 23, T - 1):                                        # windows at the edges and inside
        for h in range(WIN):
            v = 1 if h == 0 else 2 if h == WIN - 2 else 3 if h == WIN - 1 else 0
            k = n + h                                                      # = (n - 17 + h) + 17
            got = variants[v][:, k, :]
            assert torch.allclose(got, ref[n, :, h, :], atol=2e-5), (n, h)


def test_row_index_is_frame_plus_34_utterance_plus_row():
    """The table row of (utterance u, window g_rel, row h) is frame_offs[u] + g_rel + 34 u + h: rows of different
    utterances never collide and every (window, row) of a batch maps to exactly one table row."""
    T = [3, 0, 40, 1, 17]
    offs = np.concatenate([[0], np.cumsum(T)])
    seen = {}
    for u, t in enumerate(T):
        for g in range(t):
            for h in range(WIN):
                row = offs[u] + g + 34 * u + h
                key = (u, g + h)                                           # (utterance, virtual frame + 17)
                assert seen.setdefault(row, key) == key
    # an utterance owns the half-open range [offs[u] + 34 u, offs[u + 1] + 34 (u + 1))
    for row, (u, k) in seen.items():
        assert offs[u] + 34 * u <= row < offs[u + 1] + 34 * (u + 1)
        assert k == row - (offs[u] + 34 * u)

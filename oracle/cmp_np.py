"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's `cmp` composition.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
(hts-train-world_b200/) never does.

Follows, line by line in meaning:
  * data/scripts/window.pl:55-135  — delta windows over one stream of float32 statics
  * data/Makefile.in:276-321       — streams merged side by side (`merge +f -s 0 -l .. -L ..`:
                                     the file argument goes in front of stdin's frame)
  * data/scripts/addhtkheader.pl   — the 12-byte HTK header

Pinned by tests/test_cmp.py against tests/golden/cmp_perl.npz, which was written by RUNNING the
reference's own perl scripts (tests/golden/make_golden_cmp.py).
"""
import struct

import numpy as np

IGNORE = -1.0e+10   # window.pl:55


def window_stream(static, windows):
    """static: [T, dim] float32 of ONE utterance; windows: sequence of coefficient sequences.
    Returns [T, len(windows) * dim] float32 laid out as window.pl:121 does
    (t * nwin * dim + dim * (i - 1) + j)."""
    static = np.asarray(static, np.float32)
    T, dim = static.shape
    orig = static.astype(np.float64)                     # unpack("f") widens to a perl double
    out = np.zeros((T, len(windows) * dim), np.float64)
    for i, win in enumerate(windows):
        win = [float(v) for v in win]
        size = len(win)
        if size % 2 != 1:
            raise ValueError("Size of window must be 2*n + 1")           # window.pl:83-85
        nlr = (size - 1) // 2
        chk = [True] * size                                               # window.pl:70-81
        for j in range(size):
            if win[j] != 0.0:
                break
            chk[j] = False
        for j in range(size - 1, -1, -1):
            if win[j] != 0.0:
                break
            chk[j] = False
        acc = np.zeros((T, dim))
        boundary = np.zeros((T, dim), bool)
        t = np.arange(T)
        for k in range(-nlr, nlr + 1):                                    # same order as window.pl:108
            l = np.clip(t + k, 0, T - 1)
            v = orig[l]
            if chk[k + nlr]:
                boundary |= v == IGNORE
            acc = acc + win[k + nlr] * v                                  # rounded product, then add
        out[:, i * dim:(i + 1) * dim] = np.where(boundary, IGNORE, acc)
    return out.astype(np.float32)                                         # pack("f")


def compose_cmp(streams, windows=None):
    """streams: list of [T, dim] statics of one utterance, in the Makefile's order
    (mgc, lf0, bap, vib).  windows: per stream; default static / delta / delta-delta."""
    default = ((1.0,), (-0.5, 0.0, 0.5), (1.0, -2.0, 1.0))
    cols = [window_stream(s, default if windows is None else windows[i]) for i, s in enumerate(streams)]
    return np.concatenate(cols, axis=1)


def htk_header(n_frames, samp_freq, frame_shift, byte_per_frame, kind=9):
    """addhtkheader.pl: pack("l") pack("l") pack("s") pack("s"), native byte order."""
    return struct.pack("=iihh", int(n_frames), int(10000000 * frame_shift / samp_freq), int(byte_per_frame), int(kind))

"""Composition of the training observation vectors (data/Makefile.in:276-321): delta windows
(data/scripts/window.pl), stream merge, HTK header (data/scripts/addhtkheader.pl).

CPU: the numpy restatement (oracle/cmp_np.py) is pinned BIT-EXACTLY to tests/golden/cmp_perl.npz,
which was written by the reference's own perl scripts (tests/golden/make_golden_cmp.py).
GPU: the CUDA composer (wb_cmp.cu through the C ABI) must equal both, bit for bit (float32
outputs of integer-like bookkeeping plus a fixed-order double accumulation)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import cmp_np

STREAMS = ("mgc", "lf0", "bap", "vib")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "cmp_perl.npz")))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_oracle_matches_window_pl(g):
    for name in STREAMS:
        for u in range(len(g["lengths"])):
            got = cmp_np.window_stream(g["%s_static_%d" % (name, u)], ((1.0,), (-0.5, 0.0, 0.5), (1.0, -2.0, 1.0)))
            assert np.array_equal(_bits(got), _bits(g["%s_windowed_%d" % (name, u)])), (name, u)


def test_oracle_zero_taps_are_not_boundary_checked(g):
    got = cmp_np.window_stream(g["lf0_static_3"], (tuple(g["odd_window"]),))
    assert np.array_equal(_bits(got), _bits(g["odd_windowed"]))
    assert (got == cmp_np.IGNORE).any() and (got != cmp_np.IGNORE).any()


def test_oracle_htk_header(g):
    n, samp, shift, byte, kind = (int(v) for v in g["htk_args"])
    assert cmp_np.htk_header(n, samp, shift, byte, kind) == g["htk_header"].tobytes()


def test_host_htk_header_matches_perl(g):
    """wb200_htk_header is host arithmetic: runs without a GPU."""
    import hts_train_world_b200 as m
    n, samp, shift, byte, kind = (int(v) for v in g["htk_args"])
    assert m.htk_header(n, samp, shift, byte, kind) == g["htk_header"].tobytes()
    assert m.parse_window_file("3 -0.5 0.0 0.5\n") == (-0.5, 0.0, 0.5)


# ---- GPU ---------------------------------------------------------------------------------------
def _corpus(wb, lengths, fs=48000):
    # utterance u gets exactly lengths[u] frames: GetSamplesForDIO = int(1000 * n / fs / 5) + 1
    x_len = [max(1, (T - 1) * (fs // 200) + 1) for T in lengths]
    c = wb.Corpus(fs, x_len)
    assert list(c.f_len) == list(lengths)
    return c


@pytest.mark.gpu
def test_gpu_compose_matches_perl_and_oracle(wb, g):
    lengths = [int(v) for v in g["lengths"]]
    c = _corpus(wb, lengths)
    statics = {n: np.concatenate([g["%s_static_%d" % (n, u)] for u in range(len(lengths))]) for n in STREAMS}
    cmp = c.compose_cmp([(n, statics[n]) for n in STREAMS])
    assert cmp.shape == (sum(lengths), 3 * (50 + 2 + 25 + 2)) and c.cmp_dim == cmp.shape[1]
    o = 0
    for u, T in enumerate(lengths):
        want = np.concatenate([g["%s_windowed_%d" % (n, u)] for n in STREAMS], axis=1)      # merge: side by side
        assert np.array_equal(_bits(cmp[o:o + T]), _bits(want)), u
        assert np.array_equal(_bits(cmp[o:o + T]),
                              _bits(cmp_np.compose_cmp([g["%s_static_%d" % (n, u)] for n in STREAMS]))), u
        o += T
    # zero leading / trailing taps
    c2 = _corpus(wb, [lengths[3]])
    got = c2.compose_cmp([("lf0x", g["lf0_static_3"])], windows=[(tuple(g["odd_window"]),)])
    assert np.array_equal(_bits(got), _bits(g["odd_windowed"]))
    # column statistics = the sums the corpus-level reduce needs
    st = c.cmp_stats()
    ref = cmp.astype(np.float64)
    assert np.array_equal(st[:, 0], np.full(cmp.shape[1], cmp.shape[0], float))
    keep = ~(ref == cmp_np.IGNORE).any(axis=0)
    assert np.allclose(st[keep, 1], ref.sum(axis=0)[keep], rtol=1e-12, atol=1e-9)
    assert np.allclose(st[keep, 2], (ref * ref).sum(axis=0)[keep], rtol=1e-12)


@pytest.mark.gpu
def test_gpu_compose_from_the_batch_features(wb):
    """mgc / lf0 / bap taken from the batch (analysis -> code -> compose, nothing leaves the device
    in between) equal the oracle applied to the downloaded float32 features."""
    from hts_train_world_b200 import signals
    fs = 16000
    pcm = [signals.make_utterance(s, fs, duration=d)[0].numpy() for s, d in ((3, 0.6), (4, 0.9))]
    c = wb.Corpus(fs, [len(p) for p in pcm])
    c.upload_pcm16(np.concatenate(pcm))
    c.analyze()
    c.code(mgc_dim=50, bap_dim=24)
    lf0, mgc, bap = c.coded()
    cmp = c.compose_cmp()
    assert cmp.shape[1] == 3 * (50 + 1 + 24)
    for u in range(2):
        sl = c.frames_of(u)
        want = cmp_np.compose_cmp([mgc[sl], lf0[sl].reshape(-1, 1), bap[sl]])
        assert np.array_equal(_bits(cmp[sl]), _bits(want))
    with pytest.raises(wb.WorldB200Error):
        c.compose_cmp(windows=[((1.0, 1.0),)] * 3)          # even size: window.pl dies


@pytest.mark.gpu
def test_gpu_compose_empty_batch(wb):
    c = wb.Corpus(48000, [])
    assert c.compose_cmp([("s", np.zeros((0, 4), np.float32))]).shape == (0, 12)

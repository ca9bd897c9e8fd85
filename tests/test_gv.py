"""Global-variance statistics (scripts/Training.pl make_data_gv :1402-1456,
data/Makefile.in:447-458).  CPU: the numpy restatement of `vstat -d -o 2` (oracle/gv_np.py) against a
direct two-pass evaluation, and the host-side merge of per-batch partials.  GPU: the segmented
per-utterance variance kernel against the restatement."""
import numpy as np
import pytest

from oracle import gv_np


def _features(rng, n_frames, mgc_dim=50, bap_dim=24, voiced=0.6):
    mgc = (rng.standard_normal((n_frames, mgc_dim)) * np.linspace(2.0, 0.01, mgc_dim) + np.linspace(5, 0, mgc_dim)).astype(np.float32)
    bap = (rng.standard_normal((n_frames, bap_dim)) * 0.3 - 2.0).astype(np.float32)
    lf0 = np.where(rng.uniform(size=n_frames) < voiced, np.log(rng.uniform(80, 300, n_frames)), 0.0).astype(np.float32)
    return mgc, lf0, bap


def test_vstat_restatement_against_two_pass_variance():
    rng = np.random.default_rng(1)
    mgc, lf0, bap = _features(rng, 700)
    v = gv_np.utterance_gv(mgc, lf0, bap)
    x = mgc.astype(np.float64)
    assert np.allclose(v[:50], x.var(axis=0), rtol=1e-9, atol=1e-12)          # population variance, E[x^2] - mean^2
    lv = lf0[lf0 != 0].astype(np.float64)
    assert np.isclose(v[50], lv.var(), rtol=1e-7)
    assert np.allclose(v[51:], bap.astype(np.float64).var(axis=0), rtol=1e-7)
    # a stream without a single frame (no voiced frame): vstat prints nothing -> NaN
    assert np.isnan(gv_np.utterance_gv(mgc, np.zeros(700, np.float32), bap)[50])


def test_merge_of_partials_equals_the_variance_of_the_variances():
    from hts_train_world_b200 import corpus
    rng = np.random.default_rng(2)
    per = np.stack([gv_np.utterance_gv(*_features(rng, int(rng.integers(200, 900)))) for _ in range(12)])
    per[3, 50] = np.nan                                         # an utterance without voiced frames
    mean, var = gv_np.corpus_gv(per)
    parts = []
    for chunk in (per[:5], per[5:9], per[9:]):                  # three batches / ranks
        p = np.zeros((per.shape[1], 3))
        for c in range(per.shape[1]):
            v = chunk[:, c]
            v = v[~np.isnan(v)].astype(np.float32).astype(np.float64)
            p[c] = [len(v), v.sum(), (v * v).sum()]
        parts.append(p)
    m2, v2 = corpus.merge_gv(parts)
    assert np.allclose(m2, mean, rtol=1e-12) and np.allclose(v2, var, rtol=1e-7, atol=1e-18)


@pytest.mark.gpu
def test_gv_kernel_against_the_restatement(wb, reference_lib):
    """Ragged batch: the per-utterance variances of the coded features the batch holds (float32 lf0 /
    mgc / bap of the analysis tool) and the partials over the batch."""
    from hts_train_world_b200 import signals
    fs = 48000
    pcms = [signals.make_utterance(90 + i, fs, duration=d)[0].numpy() for i, d in enumerate([0.9, 0.4, 1.3])]
    pcms.append(np.zeros(int(0.3 * fs), np.int16))              # digital silence: no voiced frame
    c = wb.Corpus(fs, [len(p) for p in pcms])
    c.upload_pcm16(np.concatenate(pcms))
    c.analyze()
    c.code(50, 24)
    lf0, mgc, bap = c.coded()
    per, part = c.gv_stats()
    want = np.stack([gv_np.utterance_gv(mgc[c.frames_of(u)], lf0[c.frames_of(u)], bap[c.frames_of(u)]) for u in range(len(pcms))])
    assert np.isnan(per[3, 50]) and np.isnan(want[3, 50])
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(per), np.isnan(want))
    assert np.allclose(per[ok], want[ok], rtol=1e-12, atol=1e-15)     # same additions in the same order
    from hts_train_world_b200 import corpus
    mean, var = corpus.merge_gv([part])
    m_ref, v_ref = gv_np.corpus_gv(want)
    assert np.allclose(mean, m_ref, rtol=1e-9) and np.allclose(var, v_ref, rtol=1e-6, atol=1e-16)
    c.close()

"""Coded-aperiodicity decode of the synth tool (W/test/synth.cpp:221-247 with the SPTK code of
W/test/sptkfunctions.cpp:186-275).  CPU: the numpy restatement (oracle/bap_np.py) against the reference's own
compiled mgc2sp.  GPU: wb200_batch_set_coded_f32 against the restatement, and a Synthesis run from coded files."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import bap_np


@pytest.fixture(scope="module")
def sptk_ref():
    p = os.path.join(ROOT, "oracle", "_ref", "libsptk_ref.so")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference/externs/WORLD_v2/test"):
            import subprocess
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libsptk_ref.so"], check=True, capture_output=True)
        else:
            pytest.skip("oracle/_ref/libsptk_ref.so not built and /root/reference is absent")
    return p


@pytest.mark.parametrize("m,fft_size", [(24, 2048), (24, 1024), (12, 512), (34, 4096)])
def test_restatement_equals_the_compiled_mgc2sp(sptk_ref, m, fft_size):
    rng = np.random.default_rng(m + fft_size)
    for _ in range(3):
        coded = np.concatenate([[rng.uniform(-3, 1) - bap_np.C0_SHIFT], rng.standard_normal(m) * 0.4])   # m + 1 coefficients
        x, ap = bap_np.decode_row(coded, fft_size)
        x_ref = bap_np.reference_mgc2sp(coded, fft_size)
        assert np.max(np.abs(x - x_ref[:fft_size // 2 + 1])) <= 1e-12 * max(1.0, np.max(np.abs(x_ref)))
        # the tool's row: exp(x[j]) / 1e4 for j < m (W/test/synth.cpp:243-245)
        assert np.allclose(ap[:m], np.exp(x_ref[:m]) / 1e4, rtol=1e-12)
        # the matrix form the kernel uses
        c = coded.copy()
        c[0] += bap_np.C0_SHIFT
        assert np.max(np.abs(bap_np.decode_matrix(m, fft_size) @ c - x)) <= 1e-12 * max(1.0, np.max(np.abs(x)))


@pytest.mark.gpu
@pytest.mark.parametrize("bap_dim", [25, 24])
def test_set_coded_f32_against_the_restatement(wb, reference_lib, bap_dim):
    """lf0 / mgc / bap float32 'files' of two utterances -> f0, sp, ap in the batch.  bap_dim 25 is the fully
    defined case of the tool (odd: order 24, all 25 values read); with 24 the tool reads one value past its
    buffer -- taken as 0 here, which is what the restatement is given."""
    from conftest import load_golden
    from oracle import metrics as M
    gs = [load_golden(n) for n in ("synthetic16k_u11", "arctic_a0001")]
    fs = 16000
    refs = [reference_lib.analyze(g["pcm"].astype(np.float64) / 32768.0, fs) for g in gs]
    fft = refs[0]["fft_size"]
    feats = [reference_lib.tool_features(r["f0"], r["sp"], r["ap"], fs, fft, 50, 24) for r in refs]
    lf0 = np.concatenate([f[0] for f in feats]).astype(np.float32)
    mgc = np.concatenate([f[1] for f in feats]).astype(np.float32)
    rng = np.random.default_rng(3)
    F = len(lf0)
    bap = np.concatenate([(rng.uniform(-2, 0.5, (F, 1)) - bap_np.C0_SHIFT), rng.standard_normal((F, bap_dim - 1)) * 0.2], axis=1).astype(np.float32)
    c = wb.Corpus(fs, [len(g["pcm"]) for g in gs])
    assert c.total_frames == F
    c.set_coded_f32(fft, lf0, mgc, bap)
    f0 = c.f0()
    want_f0 = np.where(lf0 != 0, np.exp(lf0.astype(np.float64)), 0.0)
    assert np.allclose(f0, want_f0, rtol=1e-14)
    m = bap_dim - 1 if bap_dim % 2 else bap_dim
    ap = c.ap()
    rows = rng.choice(F, 40, replace=False)
    for r in rows:
        coded = np.zeros(m + 1)
        coded[:min(bap_dim, m + 1)] = bap[r, :min(bap_dim, m + 1)].astype(np.float64)
        _, want = bap_np.decode_row(coded, fft)
        assert np.allclose(ap[r], want, rtol=1e-10)
    mm = mgc.astype(np.float64).copy()
    mm[:, 0] -= 12.0
    sp_ref = reference_lib.decode_spectral_envelope(mm, fs, fft) * 1e-4
    assert M.lsd_db(sp_ref, c.sp())[1] <= 1e-4
    c.synthesis()                                   # Synthesis-only run from coded files (config 4): finite output
    assert np.isfinite(c.y()).all()
    c.close()

"""GPU: the reference's OWN command-line tools (externs/WORLD_v2/test/analysis.cpp, synth.cpp —
compiled unmodified by oracle/Makefile `tools`) linked against libworld_b200.so instead of
libworld.a, run next to the same tools linked against the reference library.  This is the
drop-in claim of BASELINE.json's north_star at the level data/Makefile.in:214 uses it."""
import os
import subprocess
import wave

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import metrics as M

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref")


def _tool(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.skip("%s not built (make -C oracle tools, needs /root/reference)" % p)
    return p


def _write_wav(path, pcm, fs):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(fs)
        w.writeframes(pcm.astype("<i2").tobytes())


def _read_wav(path):
    with wave.open(path, "rb") as w:
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float64)


def _run(args):
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["arctic_a0001", "synthetic48k_u7"])
def test_analysis_tool_is_a_drop_in(tmp_path, name):
    g = load_golden(name)
    fs, fft = int(g["fs"]), int(g["fft_size"])
    wav = str(tmp_path / "in.wav")
    _write_wav(wav, g["pcm"], fs)
    out = {}
    for tag in ("ref", "b200"):
        files = [str(tmp_path / ("%s.%s" % (tag, e))) for e in ("lf0", "mgc", "bap")]
        # data/Makefile.in:214: analysis wav lf0 mgc bap 5 FFTLEN MGCDIM  (bap dimension defaults to 24)
        _run([_tool("analysis_" + tag), wav] + files + ["5", str(fft), "50"])
        out[tag] = [np.fromfile(f, np.float32).astype(np.float64) for f in files]
    lf0_r, mgc_r, bap_r = out["ref"]
    lf0_b, mgc_b, bap_b = out["b200"]
    assert len(lf0_r) == len(lf0_b) == len(g["f0"]) and len(mgc_r) == len(mgc_b) == 50 * len(lf0_r)
    assert len(bap_r) == len(bap_b) == 24 * len(lf0_r)
    f0_r, f0_b = np.where(lf0_r != 0, np.exp(lf0_r), 0.0), np.where(lf0_b != 0, np.exp(lf0_b), 0.0)
    assert M.vuv_agreement(f0_r, f0_b) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(f0_r, f0_b) <= M.TOL_F0_REL
    same = (f0_r > 0) == (f0_b > 0)
    rows = np.repeat(same, 50)
    # 0.01 dB of log spectral distance is 1.15e-3 nepers per bin; the orthonormal DCT keeps that norm
    assert np.max(np.abs(mgc_r[rows] - mgc_b[rows])) <= 1.15e-3 * np.sqrt(fft / 2)
    assert np.max(np.abs(bap_r[np.repeat(same, 24)] - bap_b[np.repeat(same, 24)])) <= 1.15e-3 * np.sqrt(fft / 2)


def test_synth_tool_is_a_drop_in(tmp_path):
    g = load_golden("synthetic16k_u11")
    fs, fft = int(g["fs"]), int(g["fft_size"])
    wav = str(tmp_path / "in.wav")
    _write_wav(wav, g["pcm"], fs)
    raw = [str(tmp_path / ("ref.%s" % e)) for e in ("f0", "sp", "ap")]
    _run([_tool("analysis_ref"), wav] + raw + ["5", str(fft)])          # uncompressed float32 f0 / sp / ap
    ys = {}
    for tag in ("ref", "b200"):
        o = str(tmp_path / ("%s.wav" % tag))
        _run([_tool("synth_" + tag)] + raw + [o, "5", str(fft), str(fs)])
        ys[tag] = _read_wav(o)
    assert len(ys["ref"]) == len(ys["b200"]) > 0
    assert np.max(np.abs(ys["ref"] - ys["b200"])) <= 1.0            # 16-bit truncation of values 1e-6 apart
    assert M.snr_db(ys["ref"], ys["b200"]) >= M.TOL_SNR_DB


def test_corpus_driver_writes_the_tools_files(tmp_path):
    """hts-train-world_b200/driver.py = the `features:` loop of data/Makefile.in:121-242: same
    float32 lf0 / mgc / bap files as the reference tool run file by file, clip check included."""
    from hts_train_world_b200 import driver, signals
    import hts_train_world_b200 as wb
    wb.init(0)
    fs = 16000
    raw_dir = tmp_path / "raw"
    raw_dir.mkdir()
    names = []
    for i, d in enumerate([0.6, 1.1, 0.4]):
        pcm = signals.make_utterance(80 + i, fs, duration=d)[0].numpy()
        pcm.astype("<i2").tofile(str(raw_dir / ("utt%d.raw" % i)))
        names.append(("utt%d" % i, pcm))
    clipped = names[0][1].copy()
    clipped[100] = 32767
    clipped.astype("<i2").tofile(str(raw_dir / "clipped.raw"))
    (raw_dir / "empty.raw").write_bytes(b"")
    paths = sorted(str(p) for p in raw_dir.glob("*.raw"))
    # batches of ~1 s: three batches go through the reader / GPU / writer pipeline with alternating buffer sets
    rep = driver.extract_features(driver.raw_source(paths), str(tmp_path), fs=fs, log=lambda *_: None, batch_seconds=1.0)
    assert sorted(rep["skipped"]) == ["clipped", "empty"] and sorted(rep["done"]) == ["utt0", "utt1", "utt2"]
    assert not (tmp_path / "lf0" / "clipped.lf0").exists()
    n_voiced = 0
    for base, pcm in names:
        wav = str(tmp_path / (base + ".wav"))
        _write_wav(wav, pcm, fs)
        ref = [str(tmp_path / ("ref_%s.%s" % (base, e))) for e in ("lf0", "mgc", "bap")]
        _run([_tool("analysis_ref"), wav] + ref + ["5", "1024", "50"])
        lf0_r, mgc_r, bap_r = (np.fromfile(f, np.float32).astype(np.float64) for f in ref)
        lf0_b = np.fromfile(str(tmp_path / "lf0" / (base + ".lf0")), np.float32).astype(np.float64)
        mgc_b = np.fromfile(str(tmp_path / "mgc" / (base + ".mgc")), np.float32).astype(np.float64)
        bap_b = np.fromfile(str(tmp_path / "bap" / (base + ".bap")), np.float32).astype(np.float64)
        assert len(lf0_b) == len(lf0_r) and len(mgc_b) == len(mgc_r) and len(bap_b) == len(bap_r)
        same = (lf0_r != 0) == (lf0_b != 0)
        assert same.mean() >= 0.99
        assert np.max(np.abs(np.exp(lf0_b[same & (lf0_r != 0)]) / np.exp(lf0_r[same & (lf0_r != 0)]) - 1)) <= M.TOL_F0_REL
        assert np.max(np.abs(mgc_r[np.repeat(same, 50)] - mgc_b[np.repeat(same, 50)])) <= 1.15e-3 * np.sqrt(512)
        assert np.max(np.abs(bap_r[np.repeat(same, 24)] - bap_b[np.repeat(same, 24)])) <= 1.15e-3 * np.sqrt(512)
        n_voiced += int((lf0_b != 0).sum())
    assert rep["stats"][0, 0] == n_voiced
    # --resume: nothing is recomputed, and the statistics of the existing files equal the device's partials
    mtime = os.path.getmtime(str(tmp_path / "mgc" / "utt1.mgc"))
    rep2 = driver.extract_features(driver.raw_source(paths), str(tmp_path), fs=fs, log=lambda *_: None, resume=True)
    assert sorted(rep2["resumed"]) == ["utt0", "utt1", "utt2"] and rep2["done"] == []
    assert os.path.getmtime(str(tmp_path / "mgc" / "utt1.mgc")) == mtime
    s2, g2 = driver.stats_from_files(str(tmp_path), rep2["resumed"], 50, 24)
    assert np.array_equal(s2[:, 0], rep["stats"][:, 0]) and np.allclose(s2[:, 1:], rep["stats"][:, 1:], rtol=1e-9, atol=1e-9)
    assert np.array_equal(g2[:, 0], rep["gv"][:, 0]) and np.allclose(g2[:, 1:], rep["gv"][:, 1:], rtol=1e-5)


@pytest.mark.gpu
def test_corpus_driver_composes_cmp_files(tmp_path):
    """`cmp:` target of data/Makefile.in:244-321 inside the driver: cmp/<base>.cmp = HTK header
    (addhtkheader.pl) + [mgc | lf0 | bap] x (static, delta, delta-delta) (window.pl), equal bit for
    bit to the oracle composition of the lf0 / mgc / bap files written in the same run."""
    from hts_train_world_b200 import driver, signals
    from oracle import cmp_np
    import hts_train_world_b200 as wb
    wb.init(0)
    fs = 16000
    raw_dir = tmp_path / "raw"
    raw_dir.mkdir()
    for i, d in enumerate([0.5, 0.8]):
        signals.make_utterance(90 + i, fs, duration=d)[0].numpy().astype("<i2").tofile(str(raw_dir / ("u%d.raw" % i)))
    paths = sorted(str(p) for p in raw_dir.glob("*.raw"))
    rep = driver.extract_features(driver.raw_source(paths), str(tmp_path), fs=fs, log=lambda *_: None, cmp=True)
    total = 0
    acc = np.zeros((3 * 75, 2))
    for base in ("u0", "u1"):
        lf0 = np.fromfile(str(tmp_path / "lf0" / (base + ".lf0")), np.float32).reshape(-1, 1)
        mgc = np.fromfile(str(tmp_path / "mgc" / (base + ".mgc")), np.float32).reshape(-1, 50)
        bap = np.fromfile(str(tmp_path / "bap" / (base + ".bap")), np.float32).reshape(-1, 24)
        want = cmp_np.compose_cmp([mgc, lf0, bap])
        blob = (tmp_path / "cmp" / (base + ".cmp")).read_bytes()
        assert blob[:12] == cmp_np.htk_header(len(lf0), fs, 80, 4 * want.shape[1], 9)     # FRAMESHIFT = 80 at 16 kHz / 5 ms
        got = np.frombuffer(blob[12:], "<f4").reshape(len(lf0), -1)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        total += len(lf0)
        acc[:, 0] += want.astype(np.float64).sum(axis=0)
        acc[:, 1] += (want.astype(np.float64) ** 2).sum(axis=0)
    st = rep["cmp_stats"]
    assert st.shape == (225, 3) and np.all(st[:, 0] == total)
    assert np.allclose(st[:, 1], acc[:, 0], rtol=1e-10, atol=1e-7) and np.allclose(st[:, 2], acc[:, 1], rtol=1e-10)

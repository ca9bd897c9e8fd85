"""GPU: parity of the CUDA path (called through the C ABI) against
  (1) the committed golden vectors produced by the unmodified reference (tests/golden/),
  (2) the compiled reference itself (oracle/_ref) on seeded synthetic inputs,
with the tolerances BASELINE.json's north_star states:
  V/UV agreement >= 99.9 % of frames, voiced F0 relative error <= 1e-4, log spectral
  distance <= 0.01 dB (max over frames), aperiodicity absolute error <= 1e-4,
  resynthesis SNR >= 60 dB.
Each stage is fed the ORACLE's upstream outputs so errors do not compound;
test_end_to_end chains our own stages."""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, load_golden
from oracle import metrics as M

pytestmark = pytest.mark.gpu


def _x(g):
    return g["pcm"].astype(np.float64) / 32768.0


def test_randn_stream_bit_exact(wb):
    gold = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "randn_first_8192.npy"))
    assert np.array_equal(wb.randn_stream(8192), gold)


def test_randn_stream_bit_exact_deep(wb, reference_lib):
    from oracle import ref
    n = 300000                      # crosses many generator chunks (1024 variates each)
    assert np.array_equal(wb.randn_stream(n), ref.randn_stream(n))


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_dio_golden(wb, name):
    g = load_golden(name)
    t, f0 = wb.dio(_x(g), int(g["fs"]))
    assert np.array_equal(t, g["t"])
    assert M.vuv_agreement(g["f0_raw"], f0) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(g["f0_raw"], f0) <= M.TOL_F0_REL


@pytest.mark.parametrize("fs,speed", [(48000, 2), (48000, 6), (48000, 12), (16000, 4), (16000, 11)])
def test_dio_with_decimation(wb, reference_lib, fs, speed):
    """DioOption.speed > 1: zero-phase IIR decimation (W/src/matlabfunctions.cpp:184-210) first."""
    from hts_train_world_b200 import signals
    x = signals.pcm_to_double(signals.make_utterance(31, fs, duration=1.6)[0])
    t_ref, f_ref = reference_lib.dio(x, fs, speed=speed)
    t, f0 = wb.dio(x, fs, speed=speed)
    assert np.array_equal(t, t_ref)
    assert M.vuv_agreement(f_ref, f0) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(f_ref, f0) <= M.TOL_F0_REL


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_stonemask_golden(wb, name):
    g = load_golden(name)
    f0 = wb.stonemask(_x(g), int(g["fs"]), g["t"], g["f0_raw"])
    assert M.vuv_agreement(g["f0"], f0) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(g["f0"], f0) <= M.TOL_F0_REL


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_harvest_golden(wb, name):
    """Harvest (W/src/harvest.cpp:1223), F0 range 71-800 Hz, against the reference's output."""
    g = load_golden(name)
    t, f0 = wb.harvest(_x(g), int(g["fs"]))
    assert np.array_equal(t, g["t"])
    assert M.vuv_agreement(g["f0_harvest"], f0) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(g["f0_harvest"], f0) <= M.TOL_F0_REL


def test_harvest_batch_and_1ms(wb, reference_lib):
    """Ragged batch through the extension API, and the frame_period == 1 ms path (:1230-1235)."""
    from hts_train_world_b200 import signals
    fs = 48000
    pcms = [signals.make_utterance(40 + i, fs, duration=d)[0].numpy() for i, d in enumerate([0.8, 2.1, 0.3])]
    c = wb.Corpus(fs, [len(p) for p in pcms])
    c.upload_pcm16(np.concatenate(pcms))
    c.harvest()
    f0 = c.f0()
    agree, n = 0.0, 0
    for u, p in enumerate(pcms):
        x = p.astype(np.float64) / 32768.0
        _, fr = reference_lib.harvest(x, fs)
        sl = c.frames_of(u)
        agree += M.vuv_agreement(fr, f0[sl]) * len(fr)
        n += len(fr)
        assert M.f0_rel_error(fr, f0[sl]) <= M.TOL_F0_REL
    assert agree / n >= M.TOL_VUV_AGREEMENT
    c.close()
    x = pcms[0].astype(np.float64) / 32768.0
    _, fr = reference_lib.harvest(x, fs, frame_period=1.0)
    _, f1 = wb.harvest(x, fs, frame_period=1.0)
    assert M.vuv_agreement(fr, f1) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(fr, f1) <= M.TOL_F0_REL


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_cheaptrick_golden(wb, name):
    g = load_golden(name)
    sp = wb.cheaptrick(_x(g), int(g["fs"]), g["t"], g["f0"])
    assert sp.shape[1] == int(g["fft_size"]) // 2 + 1
    # golden rows are float32: compare with a float32-sized allowance on top of nothing
    mean, mx = M.lsd_db(g["sp_rows"].astype(np.float64), sp[g["rows"]])
    assert mx <= M.TOL_LSD_DB


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_d4c_golden(wb, name):
    g = load_golden(name)
    ap = wb.d4c(_x(g), int(g["fs"]), g["t"], g["f0"], int(g["fft_size"]))
    assert M.ap_abs_error(g["ap_rows"].astype(np.float64), ap[g["rows"]]) <= M.TOL_AP_ABS


@pytest.mark.parametrize("name", ["synthetic48k_u7", "synthetic16k_u11"])
def test_stages_against_compiled_reference(wb, reference_lib, name):
    g = load_golden(name)
    x, fs = _x(g), int(g["fs"])
    o = reference_lib.analyze(x, fs)
    assert np.array_equal(o["f0"], g["f0"])          # golden == reference run here
    sp = wb.cheaptrick(x, fs, o["t"], o["f0"])
    assert M.lsd_db(o["sp"], sp)[1] <= M.TOL_LSD_DB
    ap = wb.d4c(x, fs, o["t"], o["f0"], o["fft_size"])
    assert M.ap_abs_error(o["ap"], ap) <= M.TOL_AP_ABS
    y_ref = reference_lib.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    y = wb.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    assert M.snr_db(y_ref, y) >= M.TOL_SNR_DB
    # default D4C threshold (0.85) exercises the LoveTrain gate
    ap_ref = reference_lib.d4c(x, fs, o["t"], o["f0"], o["fft_size"], threshold=0.85)
    ap2 = wb.d4c(x, fs, o["t"], o["f0"], o["fft_size"], threshold=0.85)
    assert M.ap_abs_error(ap_ref, ap2) <= M.TOL_AP_ABS


@pytest.mark.parametrize("name", ["arctic_a0001", "vaiueo2d"])
def test_synthesis_golden(wb, reference_lib, name):
    """Synthesis needs the full sp/ap (not stored in the fixture), so they are regenerated by
    the compiled reference and checked against the fixture rows first."""
    g = load_golden(name)
    x, fs = _x(g), int(g["fs"])
    o = reference_lib.analyze(x, fs)
    assert np.allclose(o["sp"][g["rows"]], g["sp_rows"], rtol=1e-6)
    y = wb.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    assert M.snr_db(g["y"].astype(np.float64), y) >= M.TOL_SNR_DB


def test_end_to_end_chain(wb, reference_lib):
    """Our Dio -> StoneMask -> CheapTrick -> D4C -> Synthesis against the reference's chain."""
    g = load_golden("synthetic48k_u7")
    x, fs = _x(g), 48000
    o = reference_lib.analyze(x, fs)
    t, f0r = wb.dio(x, fs)
    f0 = wb.stonemask(x, fs, t, f0r)
    assert M.vuv_agreement(o["f0"], f0) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(o["f0"], f0) <= M.TOL_F0_REL
    sp = wb.cheaptrick(x, fs, t, f0)
    ap = wb.d4c(x, fs, t, f0, o["fft_size"])
    assert M.lsd_db(o["sp"], sp)[1] <= M.TOL_LSD_DB
    assert M.ap_abs_error(o["ap"], ap) <= M.TOL_AP_ABS
    y_ref = reference_lib.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    y = wb.synthesis(f0, sp, ap, o["fft_size"], 5.0, fs)
    assert M.snr_db(y_ref, y) >= M.TOL_SNR_DB


def test_batch_equals_single_calls(wb, reference_lib):
    """Ragged batch through the extension API == the reference run utterance by utterance."""
    from hts_train_world_b200 import signals
    fs = 48000
    durs = [0.31, 1.0, 0.05, 1.7, 0.62]
    pcms = [signals.make_utterance(20 + i, fs, duration=d)[0].numpy() for i, d in enumerate(durs)]
    c = wb.Corpus(fs, [len(p) for p in pcms])
    c.upload_pcm16(np.concatenate(pcms))
    c.analyze()
    f0, sp, ap = c.f0(), c.sp(), c.ap()
    c.synthesis()
    y = c.y()
    yoff, ylen = c.y_layout()
    for u, p in enumerate(pcms):
        x = p.astype(np.float64) / 32768.0
        o = reference_lib.analyze(x, fs)
        sl = c.frames_of(u)
        assert sl.stop - sl.start == len(o["f0"])
        assert M.vuv_agreement(o["f0"], f0[sl]) >= M.TOL_VUV_AGREEMENT
        assert M.f0_rel_error(o["f0"], f0[sl]) <= M.TOL_F0_REL
        if np.array_equal(o["f0"] > 0, f0[sl] > 0):
            assert M.lsd_db(o["sp"], sp[sl])[1] <= M.TOL_LSD_DB
            assert M.ap_abs_error(o["ap"], ap[sl]) <= M.TOL_AP_ABS
            if len(o["f0"]) >= 2:
                y_ref = reference_lib.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
                assert ylen[u] == len(y_ref)
                assert M.snr_db(y_ref, y[yoff[u]:yoff[u] + ylen[u]]) >= M.TOL_SNR_DB
    st = c.lf0_stats()
    v = f0[f0 > 0]
    assert st[0] == len(v) and np.isclose(st[1], np.log(v).sum()) and np.isclose(st[2], (np.log(v) ** 2).sum())
    c.close()


# ---- codec tail -----------------------------------------------------------------
@pytest.mark.parametrize("name", ["synthetic48k_u7", "synthetic16k_u11", "vaiueo2d"])
def test_codec_against_reference(wb, reference_lib, name):
    """CodeSpectralEnvelope / DecodeSpectralEnvelope (W/src/codec.cpp:266-324) through the C ABI."""
    g = load_golden(name)
    x, fs = _x(g), int(g["fs"])
    o = reference_lib.analyze(x, fs)
    for ndim in (50, 24):
        ref_c = reference_lib.code_spectral_envelope(o["sp"] * 1e4, fs, o["fft_size"], ndim)
        our_c = wb.code_spectral_envelope(o["sp"] * 1e4, fs, o["fft_size"], ndim)
        assert np.max(np.abs(our_c - ref_c)) <= 1e-9 * max(1.0, np.max(np.abs(ref_c)))
        ref_d = reference_lib.decode_spectral_envelope(ref_c, fs, o["fft_size"])
        our_d = wb.decode_spectral_envelope(ref_c, fs, o["fft_size"])
        assert M.lsd_db(ref_d, our_d)[1] <= 1e-6


def test_batch_coded_features_and_stats(wb, reference_lib):
    """float32 lf0 / mgc / bap of the analysis tool for a ragged batch, the statistics partials the
    NCCL reduce combines, and the decode -> Synthesis entry of config 4."""
    from hts_train_world_b200 import signals
    fs = 48000
    pcms = [signals.make_utterance(60 + i, fs, duration=d)[0].numpy() for i, d in enumerate([0.7, 1.3])]
    c = wb.Corpus(fs, [len(p) for p in pcms])
    c.upload_pcm16(np.concatenate(pcms))
    refs = [reference_lib.analyze(p.astype(np.float64) / 32768.0, fs) for p in pcms]
    # inject the reference's f0 / sp / ap so only the codec is under test
    c.set_f0(np.concatenate([r["f0"] for r in refs]))
    c.set_sp_ap(refs[0]["fft_size"], np.concatenate([r["sp"] for r in refs]), np.concatenate([r["ap"] for r in refs]))
    c.code(50, 24)
    lf0, mgc, bap = c.coded()
    want = [reference_lib.tool_features(r["f0"], r["sp"], r["ap"], fs, r["fft_size"]) for r in refs]
    wl, wm, wbp = (np.concatenate([w[i] for w in want]) for i in range(3))
    assert np.array_equal(lf0, wl)
    assert np.max(np.abs(mgc - wm)) <= 2e-6 * np.max(np.abs(wm))
    assert np.max(np.abs(bap - wbp)) <= 2e-6 * max(1.0, np.max(np.abs(wbp)))
    st = c.feature_stats()
    v = wl[wl != 0].astype(np.float64)
    assert st[0, 0] == len(v) and np.isclose(st[0, 1], np.log(np.concatenate([r["f0"] for r in refs])[wl != 0]).sum())
    assert np.all(st[1:, 0] == len(wl))
    assert np.allclose(st[1:, 1], mgc.astype(np.float64).sum(axis=0), rtol=1e-9)
    assert np.allclose(st[1:, 2], (mgc.astype(np.float64) ** 2).sum(axis=0), rtol=1e-9)
    # decode the float32 mgc back into the batch and compare with the reference's decode
    c.decode_mgc(refs[0]["fft_size"], mgc)
    sp_dec = c.sp()
    m = wm.astype(np.float64).copy()
    m[:, 0] -= 12.0
    ref_dec = reference_lib.decode_spectral_envelope(m, fs, refs[0]["fft_size"]) * 1e-4
    assert M.lsd_db(ref_dec, sp_dec)[1] <= 1e-4
    c.close()


# ---- edge cases ------------------------------------------------------------------------------------
def test_all_unvoiced_and_digital_silence(wb, reference_lib):
    fs = 16000
    rng = np.random.default_rng(5)
    x = np.concatenate([np.round(rng.standard_normal(6000) * 300) / 32768.0, np.zeros(4000)])
    t = np.arange(int(1000.0 * len(x) / fs / 5.0) + 1) * 5.0 / 1000.0
    f0 = np.zeros(len(t))
    sp_ref = reference_lib.cheaptrick(x, fs, t, f0)
    sp = wb.cheaptrick(x, fs, t, f0)
    # frames over digital silence are defined by the dither alone: they must match too
    assert np.isfinite(sp).all()
    assert M.lsd_db(sp_ref, sp)[1] <= M.TOL_LSD_DB
    ap = wb.d4c(x, fs, t, f0, 1024)
    assert np.array_equal(ap, reference_lib.d4c(x, fs, t, f0, 1024))      # all rows = 1 - 1e-12
    y_ref = reference_lib.synthesis(f0, sp_ref, ap, 1024, 5.0, fs)
    y = wb.synthesis(f0, sp_ref, ap, 1024, 5.0, fs)
    assert M.snr_db(y_ref, y) >= M.TOL_SNR_DB


def test_voiced_frames_over_digital_silence(wb, reference_lib):
    """D4C / StoneMask on frames whose window is all zeros: defined by the dither only."""
    g = load_golden("arctic_a0001")
    x, fs = _x(g), 16000
    f0 = g["f0"].copy()
    f0[-6:] = 120.0                                  # the recording ends in digital silence
    ap_ref = reference_lib.d4c(x, fs, g["t"], f0, 1024)
    ap = wb.d4c(x, fs, g["t"], f0, 1024)
    assert M.ap_abs_error(ap_ref, ap) <= M.TOL_AP_ABS
    sp_ref = reference_lib.cheaptrick(x, fs, g["t"], f0)
    sp = wb.cheaptrick(x, fs, g["t"], f0)
    assert M.lsd_db(sp_ref, sp)[1] <= M.TOL_LSD_DB


def test_short_inputs(wb, reference_lib):
    fs = 16000
    from hts_train_world_b200 import signals
    pcm = signals.make_utterance(3, fs, duration=0.2)[0].numpy()
    x = pcm / 32768.0
    # f0_length <= 7: the reference's FixF0Contour returns without writing f0 (W/src/dio.cpp:265)
    xs = x[:400]
    n = int(1000.0 * len(xs) / fs / 5.0) + 1
    assert n <= 7
    sentinel = np.full(n, -1.0)
    import ctypes as C
    o = wb.dio_option()
    t = np.zeros(n)
    wb.lib().Dio(xs.ctypes.data_as(C.POINTER(C.c_double)), len(xs), fs, C.byref(o),
                 t.ctypes.data_as(C.POINTER(C.c_double)), sentinel.ctypes.data_as(C.POINTER(C.c_double)))
    assert (sentinel == -1.0).all() and np.array_equal(t, np.arange(n) * 5.0 / 1000.0)
    # a 0.2 s utterance end to end
    o = reference_lib.analyze(x, fs)
    t, f0r = wb.dio(x, fs)
    assert M.vuv_agreement(o["f0_raw"], f0r) >= M.TOL_VUV_AGREEMENT
    f0 = wb.stonemask(x, fs, t, o["f0_raw"])
    assert M.f0_rel_error(o["f0"], f0) <= M.TOL_F0_REL
    assert M.lsd_db(o["sp"], wb.cheaptrick(x, fs, t, o["f0"]))[1] <= M.TOL_LSD_DB
    assert M.ap_abs_error(o["ap"], wb.d4c(x, fs, t, o["f0"], o["fft_size"])) <= M.TOL_AP_ABS


def test_empty_batch_and_long_utterance(wb, reference_lib):
    """n_utt = 0 is a no-op; a 10 s utterance (Dio's reference FFT would be 2^19 points,
    SURVEY.md 8 table) matches the reference through every stage."""
    c = wb.Corpus(48000, [])
    c.analyze()
    assert c.total_frames == 0 and c.f0().size == 0
    c.close()
    from hts_train_world_b200 import signals
    fs = 48000
    x = signals.pcm_to_double(signals.make_utterance(77, fs, duration=10.0)[0])
    o = reference_lib.analyze(x, fs)
    t, f0r = wb.dio(x, fs)
    f0 = wb.stonemask(x, fs, t, f0r)
    assert M.vuv_agreement(o["f0"], f0) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(o["f0"], f0) <= M.TOL_F0_REL
    assert M.lsd_db(o["sp"], wb.cheaptrick(x, fs, o["t"], o["f0"]))[1] <= M.TOL_LSD_DB
    assert M.ap_abs_error(o["ap"], wb.d4c(x, fs, o["t"], o["f0"], o["fft_size"])) <= M.TOL_AP_ABS
    y_ref = reference_lib.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    assert M.snr_db(y_ref, wb.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)) >= M.TOL_SNR_DB
    _, fh_ref = reference_lib.harvest(x, fs)
    _, fh = wb.harvest(x, fs)
    assert M.vuv_agreement(fh_ref, fh) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(fh_ref, fh) <= M.TOL_F0_REL


def test_f0_extremes(wb, reference_lib):
    """f0 at the StoneMask / CheapTrick / D4C floors and at 800 Hz."""
    fs = 48000
    from hts_train_world_b200 import signals
    x = signals.pcm_to_double(signals.make_utterance(9, fs, duration=0.6)[0])
    n = int(1000.0 * len(x) / fs / 5.0) + 1
    t = np.arange(n) * 5.0 / 1000.0
    f0 = np.zeros(n)
    vals = [40.0, 40.5, 47.0, 50.0, 70.0, 70.4, 71.0, 100.0, 333.3, 799.9, 800.0, 1200.0, 3999.0, 4000.0, 4001.0]
    f0[10:10 + len(vals)] = vals
    assert np.allclose(wb.stonemask(x, fs, t, f0), reference_lib.stonemask(x, fs, t, f0), rtol=1e-4)   # north_star: voiced F0 relative error <= 1e-4
    f0c = np.where(f0 > 1500, 0.0, f0)
    assert M.lsd_db(reference_lib.cheaptrick(x, fs, t, f0c), wb.cheaptrick(x, fs, t, f0c))[1] <= M.TOL_LSD_DB
    assert M.ap_abs_error(reference_lib.d4c(x, fs, t, f0c, 2048), wb.d4c(x, fs, t, f0c, 2048)) <= M.TOL_AP_ABS


def test_repeatability(wb, reference_lib):
    """Same call twice: every stage is bit-identical from run to run -- the analysis has no atomics, and
    the overlap-add of Synthesis accumulates fixed-point integers (W/src/synthesis.cpp:376-383 adds the
    responses in pulse order; an atomic floating-point add would depend on the order of arrival)."""
    g = load_golden("synthetic16k_u11")
    x, fs = _x(g), 16000
    a = wb.cheaptrick(x, fs, g["t"], g["f0"])
    b = wb.cheaptrick(x, fs, g["t"], g["f0"])
    assert np.array_equal(a, b)
    a = wb.d4c(x, fs, g["t"], g["f0"], 1024)
    b = wb.d4c(x, fs, g["t"], g["f0"], 1024)
    assert np.array_equal(a, b)
    o = reference_lib.analyze(x, fs)
    ys = [wb.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs) for _ in range(3)]
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[2])
    # a 48 kHz batch: many work items of different utterances in flight at once
    from hts_train_world_b200 import signals
    pcms = [signals.make_utterance(80 + i, 48000, duration=0.6)[0].numpy() for i in range(6)]
    c = wb.Corpus(48000, [len(p) for p in pcms])
    c.upload_pcm16(np.concatenate(pcms))
    c.analyze()
    c.synthesis()
    y1 = c.y().copy()
    c.synthesis()
    assert np.array_equal(y1, c.y())
    c.close()


def test_full_size_properties(wb):
    """At BASELINE.json's 48 kHz batch scale the oracle is too slow to run on everything, so
    check size-independent properties: an utterance analysed inside a 96-utterance batch gives
    bit-identical f0 / sp / ap to the same utterance analysed alone (independence of batching),
    unvoiced rows of ap are exactly 1 - 1e-12, sp is finite and positive."""
    from hts_train_world_b200 import signals
    fs = 48000
    pcms = signals.make_corpus(96, fs, first=100)
    pcms = [p.numpy() for p in pcms]
    c = wb.Corpus(fs, [len(p) for p in pcms])
    c.upload_pcm16(np.concatenate(pcms))
    c.analyze()
    f0, sp, ap = c.f0(), c.sp(), c.ap()
    assert np.isfinite(sp).all() and (sp > 0).all()
    assert np.all(ap[f0 == 0] == 1.0 - 1e-12)
    assert np.all((ap > 0) & (ap <= 1.0))
    for u in [0, 37, 95]:
        s = wb.Corpus(fs, [len(pcms[u])])
        s.upload_pcm16(pcms[u])
        s.analyze()
        sl = c.frames_of(u)
        assert np.array_equal(s.f0(), f0[sl])
        assert np.array_equal(s.sp(), sp[sl])
        assert np.array_equal(s.ap(), ap[sl])
        s.close()
    c.close()


@pytest.mark.gpu
def test_synthesis_from_float32_parameter_files(wb, reference_lib):
    """BASELINE config 4 entry: the synth tool's raw float32 f0 / sp / ap (W/test/synth.cpp:160-190,
    spec_dimension == 0) go through wb200_batch_set_params_f32; the waveform must match the
    reference's Synthesis fed the same widened values (SNR >= 60 dB, north_star)."""
    gs = [load_golden(n) for n in ("synthetic16k_u11", "arctic_a0001")]
    fs = int(gs[0]["fs"])
    assert all(int(g["fs"]) == fs for g in gs)
    refs = [reference_lib.analyze(_x(g), fs) for g in gs]
    fft = refs[0]["fft_size"]
    c = wb.Corpus(fs, [len(g["pcm"]) for g in gs])
    f0 = np.concatenate([r["f0"] for r in refs]).astype(np.float32)
    sp = np.concatenate([r["sp"] for r in refs]).astype(np.float32)
    ap = np.concatenate([r["ap"] for r in refs]).astype(np.float32)
    c.set_params_f32(fft, f0, sp, ap)
    c.synthesis()
    y = c.y()
    off, ln = c.y_layout()
    for u, r in enumerate(refs):
        sl = c.frames_of(u)
        want = reference_lib.synthesis(f0[sl].astype(np.float64), sp[sl].astype(np.float64), ap[sl].astype(np.float64),
                                       fft, 5.0, fs)
        got = y[off[u]:off[u] + ln[u]]
        assert len(got) == len(want)
        assert M.snr_db(want, got) >= 60.0


@pytest.mark.gpu
@pytest.mark.parametrize("deferred", [False, True])
def test_back_to_back_passes_with_asynchronous_copies(wb, deferred):
    """deferred: wb200_set_copy_deferral(1) -- the copies are only recorded and start right before D4C's main kernel
    of the next pass, or when something waits for them; the results must be the same.
    A long corpus run re-uses its batch objects without ever calling wb200_sync(): uploads on the
    upload stream, result copies on the download stream, stages on the library stream.  Three passes
    over the same batch object (different inputs in turn) must deliver, in double-buffered pinned host
    memory, exactly what the synchronous calls deliver -- i.e. a pass never overwrites a buffer that an
    in-flight copy of the previous pass still reads."""
    import torch
    from hts_train_world_b200 import signals
    fs = 16000
    sets = []
    for seed in (21, 22):
        pcm = [signals.make_utterance(seed + 10 * k, fs, duration=0.7)[0] for k in range(2)]
        sets.append(torch.cat(pcm).pin_memory())
    lengths = [len(sets[0]) // 2, len(sets[0]) - len(sets[0]) // 2]
    assert len(sets[1]) == len(sets[0])
    c = wb.Corpus(fs, lengths)
    F = c.total_frames
    want = []
    for s in sets:                                     # synchronous reference results
        c.upload_pcm16(s.numpy())
        c.analyze()
        c.code(50, 24)
        c.synthesis()
        want.append((c.coded(), c.y_pcm16().copy()))
    n_y = len(want[0][1])
    bufs = [dict(lf0=torch.empty(F, dtype=torch.float32).pin_memory(), mgc=torch.empty((F, 50), dtype=torch.float32).pin_memory(),
                 bap=torch.empty((F, 24), dtype=torch.float32).pin_memory(), y=torch.empty(n_y, dtype=torch.int16).pin_memory())
            for _ in range(2)]
    order = [0, 1, 0, 1]
    wb.set_copy_deferral(deferred)
    c.upload_pcm16_async(sets[order[0]])
    for it, k in enumerate(order):
        c.analyze()
        b = bufs[it % 2]
        c.code(50, 24)
        c.coded_async(b["lf0"], b["mgc"], b["bap"])
        c.synthesis()
        c.y_pcm16_async(b["y"])
        if it + 1 < len(order):
            c.upload_pcm16_async(sets[order[it + 1]])  # ordered behind the stages queued so far
    wb.sync()
    wb.set_copy_deferral(False)
    for it in (len(order) - 2, len(order) - 1):        # the last two passes own the two buffer sets
        b, ((lf0, mgc, bap), y) = bufs[it % 2], want[order[it]]
        assert np.array_equal(b["lf0"].numpy(), lf0), it
        assert np.array_equal(b["mgc"].numpy(), mgc), it
        assert np.array_equal(b["bap"].numpy(), bap), it
        assert np.array_equal(b["y"].numpy(), y), it       # the overlap-add is order-independent (integer sums)


# ---- sampling rates whose transform sizes / band counts no other test reaches ----------------------
# fs -> (CheapTrick N, D4C internal size, LoveTrain size, aperiodicity bands):
#    8 000 -> ( 512,  1024,  1024, 0)   run-time-size templates, no band at all (W/src/d4c.cpp:351-353)
#   24 000 -> (1024,  2048,  2048, 3)   odd band count: the last band transform carries one band
#   32 000 -> (2048,  4096,  4096, 4)
#   44 100 -> (2048,  4096,  4096, 5)
#   96 000 -> (4096,  8192,  8192, 5)   512-thread D4C launch + histogram selection, run-time LoveTrain size
@pytest.mark.parametrize("fs", [8000, 24000, 32000, 44100, 96000])
def test_other_sampling_rates(wb, reference_lib, fs):
    """Every stage against the compiled reference, each fed the reference's upstream outputs
    (W/src/d4c.cpp:344-357, W/src/cheaptrick.cpp:191-194 give the sizes above)."""
    from hts_train_world_b200 import signals
    x = signals.pcm_to_double(signals.make_utterance(50 + fs // 1000, fs, duration=0.9)[0])
    o = reference_lib.analyze(x, fs)
    t, f0r = wb.dio(x, fs)
    assert np.array_equal(t, o["t"])
    assert M.vuv_agreement(o["f0_raw"], f0r) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(o["f0_raw"], f0r) <= M.TOL_F0_REL
    f0 = wb.stonemask(x, fs, o["t"], o["f0_raw"])
    assert M.vuv_agreement(o["f0"], f0) >= M.TOL_VUV_AGREEMENT
    assert M.f0_rel_error(o["f0"], f0) <= M.TOL_F0_REL
    assert np.count_nonzero(o["f0"]) > 20
    sp = wb.cheaptrick(x, fs, o["t"], o["f0"])
    assert sp.shape == o["sp"].shape
    assert M.lsd_db(o["sp"], sp)[1] <= M.TOL_LSD_DB
    ap = wb.d4c(x, fs, o["t"], o["f0"], o["fft_size"])            # threshold 0, the analysis tool's setting
    if fs >= 16000:
        assert M.ap_abs_error(o["ap"], ap) <= M.TOL_AP_ABS
    else:
        # below 15.8 kHz the reference's LoveTrain sums uninitialised heap memory (boundary2 > fft_size / 2,
        # d4c.cpp:243-246), so WHICH voiced frames pass its gate differs from run to run; the rows of the
        # frames it did process are defined (knots {0: -60 dB, fs/2: -1e-12 dB}, no band at all) and ours
        # processes every voiced frame
        done = o["ap"][:, 0] < 0.5
        assert M.ap_abs_error(o["ap"][done], ap[done]) <= M.TOL_AP_ABS
        assert np.all(ap[o["f0"] == 0] == 1.0 - 1e-12) and np.all(ap[o["f0"] > 0][:, 0] < 0.5)
    if fs >= 16000:
        ap_ref = reference_lib.d4c(x, fs, o["t"], o["f0"], o["fft_size"], threshold=0.85)
        assert M.ap_abs_error(ap_ref, wb.d4c(x, fs, o["t"], o["f0"], o["fft_size"], threshold=0.85)) <= M.TOL_AP_ABS
    y_ref = reference_lib.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    y = wb.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    assert M.snr_db(y_ref, y) >= M.TOL_SNR_DB
    _, fh_ref = reference_lib.harvest(x, fs)
    _, fh = wb.harvest(x, fs)
    assert M.vuv_agreement(fh_ref, fh) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(fh_ref, fh) <= M.TOL_F0_REL


# ---- the inputs the precision study (tests/precision_study.py, profiles/README.md) calls decisive ----
def _hard_inputs():
    from scipy.signal import resample_poly
    from hts_train_world_b200 import signals
    g = load_golden("arctic_a0001")
    x16 = _x(g)[:32000]
    out = {"band_limited": np.round(np.clip(resample_poly(x16, 3, 1), -1, 1) * 32767.0) / 32768.0}
    x = signals.pcm_to_double(signals.make_utterance(3, 48000, duration=1.2)[0])
    out["dc_offset"] = x * 0.5 + 0.3
    out["quiet_7bit"] = np.round(x * 32768.0 / 256.0) / 32768.0
    return out


@pytest.mark.parametrize("case", ["band_limited", "dc_offset", "quiet_7bit"])
def test_hard_inputs(wb, reference_lib, case):
    """16 kHz real speech delivered at 48 kHz (the aperiodicity bands at 9-15 kHz hold quantisation
    noise 90 dB below the peak: the input on which single-precision centroid / power-spectrum
    transforms miss the tolerance), a DC offset of 0.3 and a recording at -48 dB: every stage that
    runs a single-precision transform (StoneMask, CheapTrick liftering, D4C bands, Synthesis) plus
    Dio, against the compiled reference fed the reference's upstream outputs."""
    x, fs = _hard_inputs()[case], 48000
    o = reference_lib.analyze(x, fs)
    t, f0_raw = wb.dio(x, fs)
    assert M.vuv_agreement(o["f0_raw"], f0_raw) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(o["f0_raw"], f0_raw) <= M.TOL_F0_REL
    f0 = wb.stonemask(x, fs, o["t"], o["f0_raw"])
    assert M.vuv_agreement(o["f0"], f0) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(o["f0"], f0) <= M.TOL_F0_REL
    assert M.lsd_db(o["sp"], wb.cheaptrick(x, fs, o["t"], o["f0"]))[1] <= M.TOL_LSD_DB
    assert M.ap_abs_error(o["ap"], wb.d4c(x, fs, o["t"], o["f0"], o["fft_size"])) <= M.TOL_AP_ABS
    ap_ref = reference_lib.d4c(x, fs, o["t"], o["f0"], o["fft_size"], threshold=0.85)
    assert M.ap_abs_error(ap_ref, wb.d4c(x, fs, o["t"], o["f0"], o["fft_size"], threshold=0.85)) <= M.TOL_AP_ABS
    y_ref = reference_lib.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    assert M.snr_db(y_ref, wb.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)) >= M.TOL_SNR_DB


def test_concurrent_callers(wb):
    """the replacement is thread-safe per call.  Two host threads call the drop-in CheapTrick /
    D4C / Synthesis on different utterances at the same time (ctypes releases the GIL during the calls);
    every result must be bit-identical to the same call made alone."""
    import threading
    gs = [load_golden("synthetic16k_u11"), load_golden("arctic_a0001")]
    want = []
    for g in gs:
        x, fs = _x(g), int(g["fs"])
        sp = wb.cheaptrick(x, fs, g["t"], g["f0"])
        ap = wb.d4c(x, fs, g["t"], g["f0"], int(g["fft_size"]))
        want.append((sp, ap, wb.synthesis(g["f0"], sp, ap, int(g["fft_size"]), 5.0, fs)))
    errors = []

    def worker(k):
        try:
            g = gs[k]
            x, fs = _x(g), int(g["fs"])
            for _ in range(4):
                sp = wb.cheaptrick(x, fs, g["t"], g["f0"])
                ap = wb.d4c(x, fs, g["t"], g["f0"], int(g["fft_size"]))
                y = wb.synthesis(g["f0"], sp, ap, int(g["fft_size"]), 5.0, fs)
                if not (np.array_equal(sp, want[k][0]) and np.array_equal(ap, want[k][1]) and np.array_equal(y, want[k][2])):
                    errors.append("thread %d: result differs from the single-threaded call" % k)
        except Exception as e:                      # noqa: BLE001
            errors.append("thread %d: %r" % (k, e))

    ths = [threading.Thread(target=worker, args=(k,)) for k in (0, 1, 0, 1)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors

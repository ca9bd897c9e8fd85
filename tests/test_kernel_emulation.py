"""CPU: the D4C kernels of hts-train-world_b200/csrc/wb_d4c.cu, the CheapTrick kernel of
wb_cheaptrick.cu, the StoneMask kernel of wb_stonemask.cu, the seven Synthesis kernels of
wb_synthesis.cu and the Dio kernels of wb_dio.cu / wb_zerocross.cuh, compiled for the CPU by the CUDA-on-CPU shim of tests/emu/ (every CUDA thread an OS thread, one CTA at a time), against the
golden vectors and the compiled reference.  It checks the SOURCE of the kernels -- indices, layouts,
barrier placement as far as logic goes -- without a GPU; the GPU parity tests check the binaries.

  mode 0: d4c_lovetrain_kernel + d4c_main_kernel (what the library runs)
  mode 1: d4c_gd_kernel + d4c_tail_kernel (the split draft, where the source tree has it)
  mode 2: d4c_lovetrain32_kernel (FP32 transform, TMA-staged window: the GPU default) + d4c_main_kernel
  mode 3: the split draft with the FP32 LoveTrain kernel
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import metrics as M

EMU = os.path.join(ROOT, "tests", "emu")
CUDA_INC = "/usr/local/cuda/include"
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    so = str(tmp_path_factory.mktemp("emu") / "libd4c_emu.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-DWB_HOST_EMU",
                    "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "hts-train-world_b200", "csrc"), "-x", "c++",
                    os.path.join(EMU, "d4c_emu.cpp"), "-o", so, "-lpthread"], check=True, capture_output=True)
    lib = C.CDLL(so)
    src = open(os.path.join(ROOT, "hts-train-world_b200", "csrc", "wb_d4c.cu")).read()
    return lib, "WB_D4C_HAS_SPLIT" in src


def _build(src, out, extra=()):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    return subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-DWB_HOST_EMU", "-I" + CUDA_INC,
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "hts-train-world_b200", "csrc"),
                           "-x", "c++"] + [os.path.join(EMU, f) for f in src] + ["-o", out, "-lpthread"] + list(extra),
                          capture_output=True, text=True)


def run(lib, x, fs, t, f0, fft_size, rows, mode, threshold=0.0):
    x, t, f0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, t, f0))
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.zeros((len(rows), fft_size // 2 + 1))
    ap0 = np.zeros(len(f0))
    rc = lib.emu_d4c(x.ctypes.data_as(dp), len(x), fs, t.ctypes.data_as(dp), f0.ctypes.data_as(dp), len(f0), fft_size,
                     C.c_double(threshold), mode, rows.ctypes.data_as(ip), len(rows), out.ctypes.data_as(dp),
                     ap0.ctypes.data_as(dp))
    assert rc == 0
    return out, ap0


def _x(g):
    return g["pcm"].astype(np.float64) / 32768.0


def modes(has_split):
    return [0, 1, 2, 3] if has_split else [0, 2]        # bit 0: split main kernels (not in this tree), bit 1: FP32 LoveTrain


def test_kernels_match_the_golden_rows(emu):
    lib, has_split = emu
    g = load_golden("synthetic48k_u7")
    rows = g["rows"][::6]
    for mode in modes(has_split):
        ap, _ = run(lib, _x(g), int(g["fs"]), g["t"], g["f0"], int(g["fft_size"]), rows, mode)
        assert M.ap_abs_error(g["ap_rows"][::6].astype(np.float64), ap) <= 1e-6, mode


def test_long_windows_edges_and_threshold(emu, reference_lib):
    """f0 = 75 Hz (windows longer than half the transform: no staging, the even / odd halves fold),
    f0 = 100 Hz (staged), the first and last frames (windows cross the utterance edges: clamped
    gather) and the default threshold 0.85 (LoveTrain gates the frames)."""
    lib, has_split = emu
    g = load_golden("synthetic48k_u7")
    x, fs, t, n = _x(g), int(g["fs"]), g["t"], int(g["fft_size"])
    f0 = np.where(np.arange(len(t)) % 2 == 0, 75.0, 100.0)
    rows = [0, 1, 150, 151, len(t) - 2, len(t) - 1]
    ref = reference_lib.d4c(x, fs, t, f0, n, threshold=0.0)
    for mode in modes(has_split):
        ap, _ = run(lib, x, fs, t, f0, n, rows, mode)
        assert M.ap_abs_error(ref[rows], ap) <= 1e-6, mode
    rows = list(range(100, 130, 3))
    ref = reference_lib.d4c(x, fs, t, g["f0"], n, threshold=0.85)
    for mode in modes(has_split):
        ap, _ = run(lib, x, fs, t, g["f0"], n, rows, mode, threshold=0.85)
        assert M.ap_abs_error(ref[rows], ap) <= 1e-6, mode


def test_no_data_races_under_thread_sanitizer(tmp_path, reference_lib):
    """Barrier placement: the emulated kernels (256 OS threads per CTA, __syncthreads = a barrier)
    run under ThreadSanitizer -- a missing __syncthreads is a data race between two threads of the
    CTA and is reported with the source line of the kernel.  Frames: a long window (f0 75 Hz: gather
    path, folding halves), a staged one, a normal voiced frame, the last frame (utterance edge).
    The detector is shown to be live first: with one barrier dropped (WBEMU_SKIP_BARRIER) it must
    report.  (The TMA copy is a memcpy here: proxy fences are outside what this can see.)"""
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path / "d4c_tsan")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-DWB_HOST_EMU",
                        "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "hts-train-world_b200", "csrc"), "-x", "c++",
                        os.path.join(EMU, "d4c_emu.cpp"), os.path.join(EMU, "emu_main.cpp"), "-o", exe, "-lpthread"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime: " + r.stderr[-200:])
    g = load_golden("synthetic48k_u7")
    x, t, f0 = _x(g), g["t"].astype(np.float64), g["f0"].astype(np.float64).copy()
    f0[0], f0[1], f0[-1] = 75.0, 100.0, 120.0
    rows = np.array([0, 1, int(np.nonzero(g["f0"] > 0)[0][10]), len(f0) - 1], dtype=np.int32)
    x.tofile(tmp_path / "x.f64"); t.tofile(tmp_path / "t.f64"); f0.tofile(tmp_path / "f0.f64"); rows.tofile(tmp_path / "rows.i32")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")

    def tsan(mode, threshold, skip=None):
        e = dict(env, WBEMU_SKIP_BARRIER=str(skip)) if skip is not None else env
        p = subprocess.run([exe, str(tmp_path), str(mode), str(threshold)], capture_output=True, text=True, env=e, timeout=900)
        if "unexpected memory mapping" in p.stderr:
            pytest.skip("ThreadSanitizer cannot map its shadow memory here")
        return p.returncode, p.stderr.count("WARNING: ThreadSanitizer")

    assert tsan(0, 0.0, skip=20)[1] > 0                    # the detector sees a dropped barrier
    has_split = "WB_D4C_HAS_SPLIT" in open(os.path.join(ROOT, "hts-train-world_b200", "csrc", "wb_d4c.cu")).read()
    for mode, thr in [(0, 0.0), (0, 0.85), (2, 0.85)] + ([(1, 0.0), (3, 0.85)] if has_split else []):
        rc, warnings = tsan(mode, thr)
        assert (rc, warnings) == (0, 0), (mode, thr)
        ap = np.fromfile(tmp_path / "ap.f64").reshape(len(rows), -1)
        ref = reference_lib.d4c(x, int(g["fs"]), t, f0, int(g["fft_size"]), threshold=thr)
        assert M.ap_abs_error(ref[rows], ap) <= 1e-6, (mode, thr)


def test_cheaptrick_kernel_source(tmp_path, reference_lib):
    """cheaptrick_kernel<11, 128> (FP64 power spectrum, FP32 liftering pair) against the golden rows
    and, on long windows / utterance edges / unvoiced frames, against the compiled reference; then the
    same frames under ThreadSanitizer."""
    so = str(tmp_path / "libct_emu.so")
    assert _build(["cheaptrick_emu.cpp"], so, ["-fPIC", "-shared"]).returncode == 0
    lib = C.CDLL(so)
    g = load_golden("synthetic48k_u7")
    x, fs, t, n = _x(g), int(g["fs"]), g["t"].astype(np.float64), int(g["fft_size"])

    def ct(f0, rows):
        f0 = np.ascontiguousarray(f0, dtype=np.float64)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.zeros((len(rows), n // 2 + 1))
        rc = lib.emu_cheaptrick(x.ctypes.data_as(dp), len(x), fs, t.ctypes.data_as(dp), f0.ctypes.data_as(dp), len(f0), n,
                                C.c_double(-0.15), rows.ctypes.data_as(ip), len(rows), out.ctypes.data_as(dp))
        assert rc == 0
        return out
    rows = g["rows"][::3]
    assert M.lsd_db(g["sp_rows"][::3].astype(np.float64), ct(g["f0"], rows))[1] <= 1e-4      # float32 fixture rows
    f0 = np.where(np.arange(len(t)) % 3 == 0, 71.5, np.where(np.arange(len(t)) % 3 == 1, 0.0, 640.0))
    rows = [0, 1, 2, 150, 151, 152, len(t) - 3, len(t) - 2, len(t) - 1]
    ref = reference_lib.cheaptrick(x, fs, t, f0)
    assert M.lsd_db(ref[rows], ct(f0, rows))[1] <= 1e-4
    exe = str(tmp_path / "ct_tsan")
    if _build(["cheaptrick_emu.cpp", "emu_main.cpp"], exe, ["-g", "-fsanitize=thread", "-DEMU_CHEAPTRICK"]).returncode != 0:
        pytest.skip("no ThreadSanitizer runtime")
    x.tofile(tmp_path / "x.f64"); t.tofile(tmp_path / "t.f64"); f0.tofile(tmp_path / "f0.f64")
    np.array(rows, dtype=np.int32).tofile(tmp_path / "rows.i32")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    p = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=env, timeout=900)
    if "unexpected memory mapping" in p.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory here")
    assert (p.returncode, p.stderr.count("WARNING: ThreadSanitizer")) == (0, 0), p.stderr[:2000]
    assert M.lsd_db(ref[rows], np.fromfile(tmp_path / "sp.f64").reshape(len(rows), -1))[1] <= 1e-4
    p = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=dict(env, WBEMU_SKIP_BARRIER="12"), timeout=900)
    assert p.stderr.count("WARNING: ThreadSanitizer") > 0          # the detector is live


@pytest.mark.parametrize("name", ["synthetic48k_u7", "arctic_a0001", "vaiueo2d"])
def test_stonemask_kernel_source(tmp_path, name):
    """stonemask_kernel (run-time transform sizes, FP32 spectra) on the golden raw F0 of three
    sampling rates: voicing identical, refined F0 within 2e-6; the 48 kHz case again under
    ThreadSanitizer."""
    so = str(tmp_path / "libsm_emu.so")
    assert _build(["stonemask_emu.cpp"], so, ["-fPIC", "-shared"]).returncode == 0
    lib = C.CDLL(so)
    g = load_golden(name)
    x, fs, t, f0r = _x(g), int(g["fs"]), g["t"].astype(np.float64), g["f0_raw"].astype(np.float64)
    rows = np.arange(0, len(t), 3, dtype=np.int32)
    out = np.zeros(len(rows))
    assert lib.emu_stonemask(x.ctypes.data_as(dp), len(x), fs, t.ctypes.data_as(dp), f0r.ctypes.data_as(dp), len(t),
                             rows.ctypes.data_as(ip), len(rows), out.ctypes.data_as(dp)) == 0
    ref = g["f0"][rows]
    assert M.vuv_agreement(ref, out) == 1.0 and M.f0_rel_error(ref, out) <= 5e-6     # FP32 transform, twiddles by squaring
    # the default path: direct FP64 evaluation of the harmonic bins, one warp per frame (at 22.05 kHz every
    # other frame position times fs is a half-integer: the exact-index path)
    out2 = np.zeros(len(rows))
    assert lib.emu_stonemask_dft(x.ctypes.data_as(dp), len(x), fs, t.ctypes.data_as(dp), f0r.ctypes.data_as(dp), len(t),
                                 rows.ctypes.data_as(ip), len(rows), out2.ctypes.data_as(dp)) == 0
    assert M.vuv_agreement(ref, out2) == 1.0 and M.f0_rel_error(ref, out2) <= 1e-9
    if name != "synthetic48k_u7":
        return
    exe = str(tmp_path / "sm_tsan")
    if _build(["stonemask_emu.cpp", "emu_main.cpp"], exe, ["-g", "-fsanitize=thread", "-DEMU_STONEMASK"]).returncode != 0:
        pytest.skip("no ThreadSanitizer runtime")
    x.tofile(tmp_path / "x.f64"); t.tofile(tmp_path / "t.f64"); f0r.tofile(tmp_path / "f0.f64"); rows[::4].copy().tofile(tmp_path / "rows.i32")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    p = subprocess.run([exe, str(tmp_path), str(fs)], capture_output=True, text=True, env=env, timeout=900)
    if "unexpected memory mapping" in p.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory here")
    assert (p.returncode, p.stderr.count("WARNING: ThreadSanitizer")) == (0, 0), p.stderr[:2000]
    assert M.f0_rel_error(ref[::4], np.fromfile(tmp_path / "f0_refined.f64")) <= 5e-6


def test_synthesis_kernels_source(tmp_path, reference_lib):
    """The whole Synthesis chain (pulse bound, increments, running phase, pulse count / scan / write,
    classification, synth_item_kernel<11, float2>) on 0.5 s of the 48 kHz fixture, fed the compiled
    reference's f0 / sp / ap: resynthesis SNR against the reference (the FP32 channel gives ~120 dB,
    tolerance 60), then the same run under ThreadSanitizer (the overlap-add uses atomics)."""
    so = str(tmp_path / "libsyn_emu.so")
    assert _build(["synthesis_emu.cpp"], so, ["-fPIC", "-shared"]).returncode == 0
    lib = C.CDLL(so)
    g = load_golden("synthetic48k_u7")
    fs = int(g["fs"])
    o = reference_lib.analyze(_x(g)[12000:36000], fs)
    f0, sp, ap = (np.ascontiguousarray(o[k], dtype=np.float64) for k in ("f0", "sp", "ap"))
    assert 0 < np.count_nonzero(f0) < len(f0)                       # voiced and unvoiced pulses
    n = int(o["fft_size"])
    ylen = int((len(f0) - 1) * 5.0 / 1000.0 * fs) + 1
    y_ref = reference_lib.synthesis(f0, sp, ap, n, 5.0, fs)[:ylen]
    y = np.zeros(ylen)
    assert lib.emu_synthesis(f0.ctypes.data_as(dp), len(f0), sp.ctypes.data_as(dp), ap.ctypes.data_as(dp), n, C.c_double(5.0), fs,
                             ylen, y.ctypes.data_as(dp)) == 0
    assert M.snr_db(y_ref, y) >= 100.0
    exe = str(tmp_path / "syn_tsan")
    if _build(["synthesis_emu.cpp", "emu_main.cpp"], exe, ["-g", "-fsanitize=thread", "-DEMU_SYNTHESIS"]).returncode != 0:
        pytest.skip("no ThreadSanitizer runtime")
    k = 40                                                             # 0.2 s: enough pulses of both kinds
    f0[:k].tofile(tmp_path / "f0.f64"); sp[:k].tofile(tmp_path / "sp.f64"); ap[:k].tofile(tmp_path / "ap.f64")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    p = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=env, timeout=1500)
    if "unexpected memory mapping" in p.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory here")
    assert (p.returncode, p.stderr.count("WARNING: ThreadSanitizer")) == (0, 0), p.stderr[:3000]
    yk = np.fromfile(tmp_path / "y.f64")
    assert M.snr_db(reference_lib.synthesis(f0[:k], sp[:k], ap[:k], n, 5.0, fs)[:len(yk)], yk) >= 100.0


def test_dio_kernels_source(tmp_path, reference_lib):
    """The Dio chain (mean, overlap-save band filters, zero-crossing count / scan / write with the
    one-barrier ballot compaction, candidates, best contour + FixF0Contour) on 0.3 s of the 48 kHz
    fixture against the compiled reference: voicing identical, raw F0 to 1e-12; a shorter excerpt
    again under ThreadSanitizer."""
    so = str(tmp_path / "libdio_emu.so")
    assert _build(["dio_emu.cpp"], so, ["-fPIC", "-shared"]).returncode == 0
    lib = C.CDLL(so)
    g = load_golden("synthetic48k_u7")
    fs = int(g["fs"])
    x = np.ascontiguousarray(_x(g)[9600:9600 + 14400])

    def dio(sig):
        out = np.zeros(int(1000.0 * len(sig) / fs / 5.0) + 1)
        assert lib.emu_dio(sig.ctypes.data_as(dp), len(sig), fs, C.c_double(71.0), C.c_double(800.0), C.c_double(2.0),
                           C.c_double(5.0), C.c_double(0.1), out.ctypes.data_as(dp)) == 0
        return out
    ref = reference_lib.dio(x, fs)[1]
    out = dio(x)
    assert np.count_nonzero(ref) > 10
    assert M.vuv_agreement(ref, out) == 1.0 and M.f0_rel_error(ref, out) <= 1e-12
    exe = str(tmp_path / "dio_tsan")
    if _build(["dio_emu.cpp", "emu_main.cpp"], exe, ["-g", "-fsanitize=thread", "-DEMU_DIO"]).returncode != 0:
        pytest.skip("no ThreadSanitizer runtime")
    xs = np.ascontiguousarray(x[:7200])
    xs.tofile(tmp_path / "x.f64")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    p = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=env, timeout=1800)
    if "unexpected memory mapping" in p.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory here")
    assert (p.returncode, p.stderr.count("WARNING: ThreadSanitizer")) == (0, 0), p.stderr[:3000]
    refs = reference_lib.dio(xs, fs)[1]
    outs = np.fromfile(tmp_path / "f0_raw.f64")
    assert M.vuv_agreement(refs, outs) == 1.0 and M.f0_rel_error(refs, outs) <= 1e-12


def test_harvest_refinement_kernel_source(tmp_path):
    """harvest_refine_thread_kernel (Goertzel recurrences at the <= 6 harmonic bins, window walks shared between the
    overlapped slots) on the CPU against the restatement of GetRefinedF0 in oracle/harvest_np.py (two FFTs per
    candidate, W/src/harvest.cpp:587-616): a gliding harmonic signal at the 8 kHz Harvest decimates to, two base
    candidates per 1 ms frame (one near the true F0, one at its octave or empty)."""
    from oracle import harvest_np as H
    so = str(tmp_path / "libhv_emu.so")
    r = _build(["harvest_emu.cpp"], so, ["-shared", "-fPIC"])
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    fs, n_fr, nc, max_base = 8000.0, 240, 2, 3
    rng = np.random.default_rng(5)
    n = int(fs * n_fr / 1000.0) + 1
    tt = np.arange(n) / fs
    f0_true = 120.0 + 260.0 * tt                                     # 120 -> 182 Hz
    phase = 2.0 * np.pi * np.cumsum(f0_true) / fs
    y = sum(np.sin(h * phase) / h for h in range(1, 9)) + 0.01 * rng.standard_normal(n) + 0.3
    base = np.zeros((n_fr, max_base))
    fk = np.interp(np.arange(n_fr) / 1000.0, tt, f0_true)
    base[:, 0] = fk * (1.0 + 0.01 * rng.standard_normal(n_fr))
    base[:, 1] = np.where(np.arange(n_fr) % 3 == 0, 0.0, 2.0 * fk)
    slots = 7 * nc
    cand = np.full((n_fr, slots), -1.0)
    score = np.full((n_fr, slots), -1.0)
    yc, bc = np.ascontiguousarray(y), np.ascontiguousarray(base)
    rc = lib.emu_harvest_refine(yc.ctypes.data_as(dp), len(yc), C.c_double(fs), C.c_double(71.0), C.c_double(800.0),
                                bc.ctypes.data_as(dp), n_fr, nc, max_base, cand.ctypes.data_as(dp), score.ctypes.data_as(dp))
    assert rc == 0
    want_c, want_s = H.refine_candidates(y, fs, base, nc, 71.0, 800.0)
    firm = np.abs(want_s - 2.5) > 1e-6                               # away from the acceptance threshold of the score
    assert np.array_equal((cand != 0.0)[firm], (want_c != 0.0)[firm])
    both = firm & (want_c != 0.0) & (cand != 0.0)
    assert both.sum() > 1000
    assert np.max(np.abs(cand[both] - want_c[both]) / want_c[both]) <= 1e-9
    assert np.max(np.abs(score[both] - want_s[both]) / want_s[both]) <= 1e-6


def _harvest_candidates(n_fr=700, nc=2, seed=3):
    """Refined candidates / scores as the contour logic sees them: three voiced stretches with gaps, a main track with
    0.2 % jitter, an octave track in some slots, empty slots, two short blips that FixStep2 must remove."""
    rng = np.random.default_rng(seed)
    slots = 7 * nc
    cand, score = np.zeros((n_fr, slots)), np.zeros((n_fr, slots))
    k = np.arange(n_fr)
    track = 140.0 + 40.0 * np.sin(k / 90.0)
    voiced = ((k > 40) & (k < 230)) | ((k > 260) & (k < 480)) | ((k > 520) & (k < 660)) | ((k > 240) & (k < 244)) | ((k > 500) & (k < 503))
    for s in range(slots):
        on = voiced & (rng.random(n_fr) > 0.15)
        octave = (s % 3 == 2) & (rng.random(n_fr) > 0.5)
        f = track * (1.0 + 0.002 * rng.standard_normal(n_fr)) * np.where(octave, 2.0, 1.0)
        cand[:, s] = np.where(on, f, 0.0)
        score[:, s] = np.where(on, 3.0 + 20.0 * rng.random(n_fr) * np.where(octave, 0.3, 1.0), 0.0)
    cand[300:310, :] *= 1.05                                           # a jump FixStep1 rejects
    return np.ascontiguousarray(cand), np.ascontiguousarray(score)


def test_harvest_contour_kernels_source(tmp_path):
    """The warp-cooperative contour logic (harvest_fix_b_kernel) against the one-lane version on the CPU: identical bit
    for bit after FixStep3 and after the merge (FixStep4), and free of data races under ThreadSanitizer."""
    so = str(tmp_path / "libhv_emu.so")
    r = _build(["harvest_emu.cpp"], so, ["-shared", "-fPIC"])
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    cand, score = _harvest_candidates()
    n_fr, nc = cand.shape[0], cand.shape[1] // 7
    outs = [np.zeros(n_fr) for _ in range(4)]
    rc = lib.emu_harvest_contour(cand.ctypes.data_as(dp), score.ctypes.data_as(dp), n_fr, nc, *[o.ctypes.data_as(dp) for o in outs])
    assert rc == 0
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[2], outs[3])
    assert np.count_nonzero(outs[0]) > 300 and np.count_nonzero(outs[0][236:250]) == 0      # contours found, the blip is gone
    exe = str(tmp_path / "hv_tsan")
    r = _build(["harvest_emu.cpp"], exe, ["-g", "-fsanitize=thread", "-DEMU_HARVEST_MAIN"])
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + r.stderr[-200:])
    path = str(tmp_path / "cand.bin")
    with open(path, "wb") as f:
        f.write(np.array([n_fr, nc], np.int32).tobytes())
        f.write(cand.tobytes())
        f.write(score.tobytes())
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    p = subprocess.run([exe, path], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stderr.count("WARNING: ThreadSanitizer") == 0, p.stderr[-3000:]

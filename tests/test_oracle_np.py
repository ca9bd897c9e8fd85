"""CPU: pins the numpy / C restatement (oracle/world_np.py, oracle/world_port.c) against the golden
vectors written by the compiled, unmodified reference (tests/golden/) and against the compiled
reference itself.  Differences are FFT rounding (numpy pocketfft vs the reference's Ooura FFT)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden
from oracle import metrics as M


@pytest.fixture(scope="module")
def wnp():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libworld_port.so"], check=True,
                   capture_output=True)
    from oracle import world_np
    return world_np


def _x(g):
    return g["pcm"].astype(np.float64) / 32768.0


def test_randn_stream(wnp):
    assert np.array_equal(wnp.randn_stream(8192), np.load(os.path.join(GOLDEN, "randn_first_8192.npy")))


def test_interp1_extrapolates_like_the_reference(wnp):
    x = np.array([0.0, 1.0, 3.0])
    y = np.array([0.0, 10.0, 30.0])
    assert np.allclose(wnp.interp1(x, y, [-1.0, 0.0, 0.5, 1.0, 2.0, 3.0, 4.0]), [-10, 0, 5, 10, 20, 30, 40])


@pytest.mark.parametrize("name", ["arctic_a0001", "vaiueo2d", "synthetic48k_u7", "synthetic16k_u11"])
def test_dio(wnp, name):
    """The numpy restatement of Dio (W/src/dio.cpp, speed 1) against the golden raw F0 of the compiled
    reference: every voicing decision equal, F0 to FFT rounding (numpy's FFT vs Ooura's)."""
    g = load_golden(name)
    t, f0 = wnp.dio(_x(g), int(g["fs"]))
    assert np.array_equal(t, g["t"])
    assert M.vuv_agreement(g["f0_raw"], f0) == 1.0
    assert M.f0_rel_error(g["f0_raw"], f0) <= 1e-9


def test_dio_short_input_is_left_unwritten(wnp):
    # FixF0Contour returns before writing f0 when there are <= 7 frames (W/src/dio.cpp:266); the oracle yields zeros
    t, f0 = wnp.dio(np.sin(np.arange(400) * 0.1), 16000)
    assert len(t) == 6 and not f0.any()


@pytest.mark.parametrize("name", ["vaiueo2d", "synthetic16k_u11"])
def test_stonemask_and_cheaptrick(wnp, name):
    g = load_golden(name)
    x, fs = _x(g), int(g["fs"])
    f0 = wnp.stonemask(x, fs, g["t"], g["f0_raw"])
    assert M.vuv_agreement(g["f0"], f0) == 1.0 and M.f0_rel_error(g["f0"], f0) <= 1e-6
    sp = wnp.cheaptrick(x, fs, g["t"], g["f0"], fft_size=int(g["fft_size"]))
    assert M.lsd_db(g["sp_rows"].astype(np.float64), sp[g["rows"]])[1] <= 2e-6 + 1e-5   # float32 fixture rows


def test_d4c(wnp):
    g = load_golden("vaiueo2d")
    ap = wnp.d4c(_x(g), int(g["fs"]), g["t"], g["f0"], int(g["fft_size"]), threshold=0.0)
    assert M.ap_abs_error(g["ap_rows"].astype(np.float64), ap[g["rows"]]) <= 1e-6


def test_synthesis_and_codec(wnp, reference_lib):
    g = load_golden("vaiueo2d")
    x, fs = _x(g), int(g["fs"])
    o = reference_lib.analyze(x, fs)
    y = wnp.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    assert M.snr_db(g["y"].astype(np.float64), y) >= 100.0
    ref_c = reference_lib.code_spectral_envelope(o["sp"] * 1e4, fs, o["fft_size"], 50)
    assert np.max(np.abs(wnp.code_spectral_envelope(o["sp"] * 1e4, fs, o["fft_size"], 50) - ref_c)) <= 1e-9

"""CPU: the D4C kernels of hts-train-world_b200/csrc/wb_d4c.cu, compiled for the CPU by the
CUDA-on-CPU shim of tests/emu/ (every CUDA thread an OS thread, one CTA at a time), against the
golden vectors and the compiled reference.  It checks the SOURCE of the kernels -- indices, layouts,
barrier placement as far as logic goes -- without a GPU; the GPU parity tests check the binaries.

  mode 0: d4c_lovetrain_kernel + d4c_main_kernel (what the library runs)
  mode 1: d4c_gd_kernel + d4c_tail_kernel (the split draft, where the source tree has it)
  mode 3: the same with the FP32 LoveTrain kernel
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
    return [0, 1, 3] if has_split else [0]


def test_kernels_match_the_golden_rows(emu):
    lib, has_split = emu
    g = load_golden("synthetic48k_u7")
    rows = g["rows"][::3]
    for mode in modes(has_split):
        ap, _ = run(lib, _x(g), int(g["fs"]), g["t"], g["f0"], int(g["fft_size"]), rows, mode)
        assert M.ap_abs_error(g["ap_rows"][::3].astype(np.float64), ap) <= 1e-6, mode


def test_long_windows_edges_and_threshold(emu, reference_lib):
    """f0 = 75 Hz (windows longer than half the transform: no staging, the even / odd halves fold),
    f0 = 100 Hz (staged), the first and last frames (windows cross the utterance edges: clamped
    gather) and the default threshold 0.85 (LoveTrain gates the frames)."""
    lib, has_split = emu
    g = load_golden("synthetic48k_u7")
    x, fs, t, n = _x(g), int(g["fs"]), g["t"], int(g["fft_size"])
    f0 = np.where(np.arange(len(t)) % 2 == 0, 75.0, 100.0)
    rows = [0, 1, 2, 3, 150, 151, len(t) - 2, len(t) - 1]
    ref = reference_lib.d4c(x, fs, t, f0, n, threshold=0.0)
    for mode in modes(has_split):
        ap, _ = run(lib, x, fs, t, f0, n, rows, mode)
        assert M.ap_abs_error(ref[rows], ap) <= 1e-6, mode
    rows = list(range(100, 130))
    ref = reference_lib.d4c(x, fs, t, g["f0"], n, threshold=0.85)
    for mode in modes(has_split):
        ap, _ = run(lib, x, fs, t, g["f0"], n, rows, mode, threshold=0.85)
        assert M.ap_abs_error(ref[rows], ap) <= 1e-6, mode

"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, its host-only
helpers agree with the reference's formulas, and compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for root, _, files in os.walk(inc):
        for f in files:
            src = open(os.path.join(root, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            for m in re.finditer(r"WORLD_API\s+[\w\s\*]+?\b(\w+)\s*\(", src):
                names.add(m.group(1))
    return sorted(names)


def test_headers_declare_the_world_api():
    names = declared_symbols()
    for must in ["Dio", "InitializeDioOption", "GetSamplesForDIO", "StoneMask", "CheapTrick",
                 "InitializeCheapTrickOption", "GetFFTSizeForCheapTrick", "GetF0FloorForCheapTrick",
                 "D4C", "InitializeD4COption", "Synthesis", "Harvest", "InitializeHarvestOption",
                 "GetSamplesForHarvest", "wb200_batch_create", "wb200_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    import hts_train_world_b200 as wb
    lib = C.CDLL(wb.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert missing == []


def test_option_structs_match_reference_layout():
    import hts_train_world_b200 as wb
    # W/src/world/dio.h:16-23, cheaptrick.h:16-20, d4c.h:16-18, harvest.h:16-20
    assert C.sizeof(wb.DioOption) == 48 and wb.DioOption.speed.offset == 32
    assert wb.DioOption.allowed_range.offset == 40
    assert C.sizeof(wb.CheapTrickOption) == 24 and wb.CheapTrickOption.fft_size.offset == 16
    assert C.sizeof(wb.D4COption) == 8
    assert C.sizeof(wb.HarvestOption) == 24


def test_host_helpers_follow_reference_formulas():
    import hts_train_world_b200 as wb
    L = wb.lib()
    o = wb.DioOption()
    L.InitializeDioOption(C.byref(o))          # W/src/dio.cpp:649-665
    assert (o.f0_floor, o.f0_ceil, o.channels_in_octave, o.frame_period, o.speed,
            o.allowed_range) == (71.0, 800.0, 2.0, 5.0, 1, 0.1)
    for fs, n in [(16000, 1024), (22050, 1024), (44100, 2048), (48000, 2048), (8000, 512),
                  (96000, 4096)]:
        c = wb.CheapTrickOption()
        L.InitializeCheapTrickOption(fs, C.byref(c))   # W/src/cheaptrick.cpp:191-194,230-239
        assert c.fft_size == n and c.q1 == -0.15 and c.f0_floor == 71.0
        assert L.GetF0FloorForCheapTrick(fs, n) == 3.0 * fs / (n - 3.0)
    d = wb.D4COption()
    L.InitializeD4COption(C.byref(d))
    assert d.threshold == 0.85
    h = wb.HarvestOption()
    L.InitializeHarvestOption(C.byref(h))
    assert (h.f0_floor, h.f0_ceil, h.frame_period) == (71.0, 800.0, 5.0)
    for fs, n, fp in [(48000, 144000, 5.0), (16000, 53680, 5.0), (22050, 17500, 5.0), (16000, 1, 5.0),
                      (44100, 99999, 1.0)]:
        want = int(1000.0 * n / fs / fp) + 1       # W/src/dio.cpp:638-640
        assert L.GetSamplesForDIO(fs, n, fp) == want
        assert L.GetSamplesForHarvest(fs, n, fp) == want


def test_host_helpers_agree_with_compiled_reference(reference_lib):
    import hts_train_world_b200 as wb
    L, R = wb.lib(), reference_lib.lib
    for fs in [8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000]:
        a, b = wb.CheapTrickOption(), type(reference_lib.cheaptrick_option(fs))()
        L.InitializeCheapTrickOption(fs, C.byref(a))
        R.InitializeCheapTrickOption(fs, C.byref(b))
        assert (a.q1, a.f0_floor, a.fft_size) == (b.q1, b.f0_floor, b.fft_size)
        for n in [1, 7, 1000, 53680, 480001]:
            for fp in [1.0, 5.0, 10.0]:
                assert L.GetSamplesForDIO(fs, n, fp) == R.GetSamplesForDIO(fs, n, fp)


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import hts_train_world_b200 as wb
    with pytest.raises(wb.WorldB200Error):
        wb.init(0)
    x = np.zeros(16000)
    with pytest.raises(wb.WorldB200Error):
        wb.dio(x, 16000)
    with pytest.raises(wb.WorldB200Error):
        wb.cheaptrick(x, 16000, np.arange(10) * 0.005, np.zeros(10))
    with pytest.raises(wb.WorldB200Error):
        wb.Corpus(16000, [16000])


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure; nothing under the package may import, link or call it."""
    pkg = os.path.join(ROOT, "hts-train-world_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s*[<\"].*oracle)|libworld_ref|libworld_port", re.M)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not pat.search(src), "%s references the oracle" % f


def test_reference_tools_link_against_the_library():
    """The reference's unmodified analysis / synth tools, linked against libworld_b200.so instead
    of libworld.a (oracle/Makefile `tools`), load and reach their own argument check."""
    import subprocess
    bin_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir("/root/reference/externs/WORLD_v2/test"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "tools"], check=True, capture_output=True)
    for tool in ("analysis_b200", "synth_b200"):
        p = os.path.join(bin_dir, tool)
        if not os.path.exists(p):
            pytest.skip(p + " not built and /root/reference is absent")
        r = subprocess.run([p], capture_output=True, text=True)
        assert "sage" in (r.stdout + r.stderr)            # "Usage: ..." / "usage" from the tool itself

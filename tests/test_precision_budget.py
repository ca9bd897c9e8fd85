"""CPU: the precision split of D4C's transforms is pinned by emulating
single-precision transforms in the numpy oracle (tests/precision_study.py).

What the kernels rely on, and what this file keeps true:
  * the band transforms (GetCoarseAperiodicity, W/src/d4c.cpp:192-223) may run in FP32: the
    aperiodicity moves by < 1e-7 on every input, four orders inside the 1e-4 tolerance;
  * the centroid and power-spectrum transforms (:90-119, :148-164) may NOT: on a recording whose
    upper bands are empty (16 kHz speech delivered at 48 kHz -- the bands at 9-15 kHz hold
    quantisation noise 90 dB below the peak) their FP32 versions break the 1e-4 tolerance, although
    the bench corpus (noise floor at -34 dB) would never show it;
  * LoveTrain's transform (:225-250) is a candidate for FP32: ap0 moves by < 1e-7;
  * the FP32 transforms of the other kernels (StoneMask's spectra, CheapTrick's liftering pair,
    Synthesis' four transforms) stay four to six orders inside their tolerances on the same input.
"""
import numpy as np
import pytest

import precision_study as PS
from precision_study import W

TOL_AP = 1e-4          # BASELINE.json north_star: aperiodicity absolute error


@pytest.fixture(scope="module")
def band_limited():
    from scipy.signal import resample_poly
    g = dict(np.load(PS.os.path.join(PS.ROOT, "tests", "golden", "arctic_a0001.npz")))
    x = g["pcm"].astype(np.float64)[:32000] / 32768.0                   # 2 s of real speech at 16 kHz
    x48 = np.round(np.clip(resample_poly(x, 3, 1), -1, 1) * 32767.0) / 32768.0
    t, f0 = PS._contour(x48, 48000)
    assert np.count_nonzero(f0) > 100
    return x48, 48000, t, f0


def _err(sig, **kw):
    x, fs, t, f0 = sig
    n = W.cheaptrick_fft_size(fs)
    return float(np.max(np.abs(PS.d4c_variant(x, fs, t, f0, n, **kw) - PS.d4c_variant(x, fs, t, f0, n))))


def test_band_transforms_tolerate_fp32(band_limited):
    assert _err(band_limited, band32=True) < 1e-7


def test_centroid_and_power_transforms_need_fp64(band_limited):
    assert _err(band_limited, centroid="c32") > TOL_AP
    assert _err(band_limited, power32=True) > TOL_AP


def test_lovetrain_transform_tolerates_fp32(band_limited):
    x, fs, t, f0 = band_limited
    n = W.cheaptrick_fft_size(fs)
    a = PS.d4c_variant(x, fs, t, f0, n, lt32=True, return_ap0=True)[1]
    b = PS.d4c_variant(x, fs, t, f0, n, return_ap0=True)[1]
    assert float(np.max(np.abs(a - b))) < 1e-7


def test_bench_corpus_alone_would_not_show_it():
    res = PS.run(quick=True, variants=[("all", dict(centroid="c32", power32=True, band32=True, lt32=True))])
    for case, r in res.items():
        assert r["all"] < 1e-5, case


def test_fp32_transforms_of_the_other_kernels_hold_on_the_hard_input(band_limited):
    x, fs, t, f0 = band_limited
    r = PS.kernels_fp32_choices(x, fs, t, f0)
    assert r["stonemask: V/UV agreement"] == 1.0
    assert r["stonemask: F0 relative error (tol 1e-4)"] < 1e-6
    assert r["cheaptrick lifter: LSD dB (tol 0.01)"] < 1e-4
    assert r["synthesis: SNR dB (tol >= 60)"] > 100.0

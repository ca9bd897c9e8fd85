"""CPU: the link-compatibility helpers of libworld_b200.so (world/common.h, matlabfunctions.h,
fft.h and the band-aperiodicity codec of codec.h -- hts-train-world_b200/csrc/wb_compat.cu)
against the same functions of the unmodified reference compiled into oracle/_ref/libworld_ref.so.

Bar: bit-exact wherever the arithmetic has one evaluation order (everything except the
transforms); 1e-12 relative to the largest bin where an FFT is involved (an ordinary radix-2
transform here, the reference's vendored split-radix code there)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
cx = C.c_double * 2


class FftPlan(C.Structure):      # W/src/world/fft.h:26-37
    _fields_ = [("n", C.c_int), ("sign", C.c_int), ("flags", C.c_uint), ("c_in", C.c_void_p),
                ("in_", C.c_void_p), ("c_out", C.c_void_p), ("out", C.c_void_p),
                ("input", C.c_void_p), ("ip", C.c_void_p), ("w", C.c_void_p)]


class ForwardRealFFT(C.Structure):   # W/src/world/common.h:18-38 (InverseRealFFT has the same layout)
    _fields_ = [("fft_size", C.c_int), ("waveform", dp), ("spectrum", C.c_void_p), ("plan", FftPlan)]


class MinimumPhaseAnalysis(C.Structure):   # W/src/world/common.h:48-55
    _fields_ = [("fft_size", C.c_int), ("log_spectrum", dp), ("minimum_phase_spectrum", C.c_void_p),
                ("cepstrum", C.c_void_p), ("inverse_fft", FftPlan), ("forward_fft", FftPlan)]


def _bind(lib):
    lib.fft_plan_dft_1d.restype = FftPlan
    lib.fft_plan_dft_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    lib.fft_plan_dft_c2r_1d.restype = FftPlan
    lib.fft_plan_dft_c2r_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
    lib.fft_plan_dft_r2c_1d.restype = FftPlan
    lib.fft_plan_dft_r2c_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
    lib.fft_execute.argtypes = [FftPlan]
    lib.fft_execute.restype = None
    lib.fft_destroy_plan.argtypes = [FftPlan]
    lib.fft_destroy_plan.restype = None
    lib.randn.restype = C.c_double
    lib.matlab_std.restype = C.c_double
    lib.matlab_std.argtypes = [dp, C.c_int]
    lib.matlab_round.argtypes = [C.c_double]
    lib.DCCorrection.argtypes = [dp, C.c_double, C.c_int, C.c_int, dp]
    lib.LinearSmoothing.argtypes = [dp, C.c_double, C.c_int, C.c_int, dp]
    lib.interp1Q.argtypes = [C.c_double, C.c_double, dp, C.c_int, dp, C.c_int, dp]
    lib.interp1.argtypes = [dp, dp, C.c_int, dp, C.c_int, dp]
    lib.histc.argtypes = [dp, C.c_int, dp, C.c_int, ip]
    lib.decimate.argtypes = [dp, C.c_int, C.c_int, dp]
    lib.fast_fftfilt.argtypes = [dp, C.c_int, dp, C.c_int, C.c_int, C.c_void_p, C.c_void_p, dp]
    for f in ("InitializeForwardRealFFT", "InitializeInverseRealFFT", "InitializeMinimumPhaseAnalysis"):
        getattr(lib, f).argtypes = [C.c_int, C.c_void_p]
    for f in ("DestroyForwardRealFFT", "DestroyInverseRealFFT", "DestroyMinimumPhaseAnalysis",
              "GetMinimumPhaseSpectrum"):
        getattr(lib, f).argtypes = [C.c_void_p]
    return lib


def P(a):
    return a.ctypes.data_as(dp)


def rows(a2d):
    arr = (dp * a2d.shape[0])()
    for i in range(a2d.shape[0]):
        arr[i] = C.cast(a2d.ctypes.data + i * a2d.strides[0], dp)
    return arr


@pytest.fixture(scope="module")
def libs(reference_lib):
    import hts_train_world_b200 as wb
    ours = _bind(C.CDLL(wb.LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW))
    ref = _bind(reference_lib.lib)
    return ours, ref


def both(libs, fn):
    return fn(libs[0]), fn(libs[1])


def test_every_global_function_of_the_reference_library_is_exported(libs):
    """`nm -g` of the reference library (its unmangled text symbols) is a subset of ours."""
    import subprocess
    import hts_train_world_b200 as wb
    from oracle import ref as oref

    def text_syms(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
        return {l.split()[2] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] in "TW"}
    theirs = {s for s in text_syms(oref.ref_path()) if not s.startswith("_")}
    assert len(theirs) > 40
    assert sorted(theirs - text_syms(wb.LIB_PATH)) == []


def test_randn_stream_and_reseed(libs):
    def run(lib):
        lib.randn_reseed()
        a = [lib.randn() for _ in range(5000)]
        lib.randn_reseed()
        return np.array(a), lib.randn()
    (a, a0), (b, b0) = both(libs, run)
    assert np.array_equal(a, b) and a0 == b0 == a[0]
    assert np.array_equal(a[:8192], np.load(os.path.join(ROOT, "tests", "golden", "randn_first_8192.npy"))[:5000])


def test_small_helpers_bit_exact(libs):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1001)
    for lib_fn in [lambda l: l.matlab_std(P(x), 1001), lambda l: l.GetSuitableFFTSize(3000),
                   lambda l: [l.matlab_round(v) for v in (-2.5, -0.49, 0.0, 0.5, 1.49, 7.5)],
                   lambda l: [l.GetSuitableFFTSize(n) for n in (1, 2, 3, 1023, 1024, 1025, 144000)]]:
        a, b = both(libs, lib_fn)
        assert a == b

    def shift_diff_nuttall(l):
        y = np.zeros(1000); d = np.zeros(1000); w = np.zeros(513)
        l.fftshift(P(x), 1000, P(y)); l.diff(P(x), 1001, P(d)); l.NuttallWindow(513, P(w))
        return np.concatenate([y, d, w])
    a, b = both(libs, shift_diff_nuttall)
    assert np.array_equal(a, b)


def test_histc_interp1_interp1q_bit_exact(libs):
    rng = np.random.default_rng(2)
    for n, m in [(2, 7), (5, 50), (300, 601), (64, 1)]:
        x = np.sort(rng.uniform(0, 3, n)); y = rng.standard_normal(n)
        # queries below, inside, on and above the knots (interp1 extrapolates linearly)
        xi = np.sort(np.concatenate([rng.uniform(-1, 4, m), x[: min(3, n)]]))[:m + 3].copy()

        def run(l):
            idx = np.zeros(len(xi), np.int32); yi = np.zeros(len(xi))
            l.histc(P(x), n, P(xi), len(xi), idx.ctypes.data_as(ip))
            l.interp1(P(x), P(y), n, P(xi), len(xi), P(yi))
            return idx, yi
        (i1, y1), (i2, y2) = both(libs, run)
        assert np.array_equal(i1, i2) and np.array_equal(y1, y2)
        assert np.array_equal(i1, np.clip(np.searchsorted(x, xi, side="right"), 1, n - 1))
    y = rng.standard_normal(400)
    xi = np.sort(rng.uniform(10.0, 10.0 + 398 * 0.37, 900))

    def runq(l):
        out = np.zeros(900)
        l.interp1Q(10.0, 0.37, P(y), 400, P(xi), 900, P(out))
        return out
    a, b = both(libs, runq)
    assert np.array_equal(a, b)


def test_dc_correction_and_linear_smoothing_bit_exact(libs):
    rng = np.random.default_rng(3)
    for fs, n, f0 in [(16000, 1024, 71.0), (48000, 2048, 220.5), (48000, 4096, 47.0), (44100, 2048, 799.0)]:
        p = rng.uniform(0.1, 4.0, n // 2 + 1 + 8)       # DCCorrection reads one bin past its range

        def run(l):
            o = p.copy(); s = np.zeros(n // 2 + 1); s2 = np.zeros(n // 2 + 1)
            l.DCCorrection(P(p), f0, fs, n, P(o))
            l.LinearSmoothing(P(p), f0 * 2.0 / 3.0, fs, n, P(s))
            l.LinearSmoothing(P(p), f0 / 2.0, fs, n, P(s2))
            return np.concatenate([o, s, s2])
        a, b = both(libs, run)
        assert np.array_equal(a, b)


def test_decimate_bit_exact(libs):
    rng = np.random.default_rng(4)
    x = rng.standard_normal(4801)
    for r in list(range(2, 13)) + [1, 13]:             # outside 2..12 the reference's filter is all zeros
        for L in (4801, 4800, 97):
            def run(l):
                y = np.full((L - 1) // r + 1 + 10, -7.0)       # the loop writes up to 9 // r values past nout
                l.decimate(P(x), L, r, P(y))
                return y
            a, b = both(libs, run)
            assert np.array_equal(a, b), (r, L)


def test_fft_plans_match_the_reference_conventions(libs):
    rng = np.random.default_rng(5)
    for n in (2, 4, 8, 64, 1024, 4096):
        xr = rng.standard_normal(n)
        xc = rng.standard_normal((n, 2))

        def run(l):
            res = []
            a = xr.copy(); spec = np.full((n, 2), 9.0)
            p = l.fft_plan_dft_r2c_1d(n, a.ctypes.data, spec.ctypes.data, 3)
            l.fft_execute(p); l.fft_destroy_plan(p)
            res.append(spec[: n // 2 + 1].copy())
            assert np.all(spec[n // 2 + 1:] == 9.0)       # bins above n/2 are not written
            out = np.zeros(n); s2 = np.ascontiguousarray(xc.copy())
            p = l.fft_plan_dft_c2r_1d(n, s2.ctypes.data, out.ctypes.data, 3)
            l.fft_execute(p); l.fft_destroy_plan(p)
            res.append(out.copy())
            for sign in (1, 2):
                cin = np.ascontiguousarray(xc.copy()); cout = np.zeros((n, 2))
                p = l.fft_plan_dft_1d(n, cin.ctypes.data, cout.ctypes.data, sign, 3)
                l.fft_execute(p); l.fft_destroy_plan(p)
                res.append(cout.copy())
            return res
        ours, ref = both(libs, run)
        for a, b in zip(ours, ref):
            assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(b)))
        # and against the definitions: r2c = DFT; c2r = n * irfft of bins 0..n/2 (imaginary parts of
        # bins 0 and n/2 ignored); c2c forward = DFT(conj x), backward = n * IDFT(conj x)
        z = xc[:, 0] + 1j * xc[:, 1]
        assert np.allclose(ours[0][:, 0] + 1j * ours[0][:, 1], np.fft.rfft(xr), atol=1e-10)
        assert ours[0][0, 1] == 0.0 and ours[0][n // 2, 1] == 0.0
        assert np.allclose(ours[1], n * np.fft.irfft(z[: n // 2 + 1], n), atol=1e-9)
        assert np.allclose(ours[2][:, 0] + 1j * ours[2][:, 1], np.fft.fft(np.conj(z)), atol=1e-9)
        assert np.allclose(ours[3][:, 0] + 1j * ours[3][:, 1], n * np.fft.ifft(np.conj(z)), atol=1e-9)


def test_minimum_phase_and_fftfilt(libs):
    rng = np.random.default_rng(6)
    n = 2048
    logsp = np.log(rng.uniform(0.01, 5.0, n // 2 + 1)) / 2.0
    x = rng.standard_normal(700); h = rng.standard_normal(301)

    def run(l):
        m = MinimumPhaseAnalysis()
        l.InitializeMinimumPhaseAnalysis(n, C.byref(m))
        assert m.fft_size == n
        for i in range(n // 2 + 1):
            m.log_spectrum[i] = logsp[i]
        l.GetMinimumPhaseSpectrum(C.byref(m))
        spec = np.ctypeslib.as_array(C.cast(m.minimum_phase_spectrum, dp), (n, 2))[: n // 2 + 1].copy()
        l.DestroyMinimumPhaseAnalysis(C.byref(m))
        f = ForwardRealFFT(); g = ForwardRealFFT()
        l.InitializeForwardRealFFT(1024, C.byref(f)); l.InitializeInverseRealFFT(1024, C.byref(g))
        y = np.zeros(1024)
        l.fast_fftfilt(P(x), 700, P(h), 301, 1024, C.byref(f), C.byref(g), P(y))
        l.DestroyForwardRealFFT(C.byref(f)); l.DestroyInverseRealFFT(C.byref(g))
        return spec, y
    (s1, y1), (s2, y2) = both(libs, run)
    assert np.max(np.abs(s1 - s2)) <= 1e-12 * np.max(np.abs(s2))
    assert np.max(np.abs(y1 - y2)) <= 1e-12 * np.max(np.abs(y2))
    # |minimum-phase spectrum| = exp(log spectrum)
    assert np.allclose(np.hypot(s1[:, 0], s1[:, 1]), np.exp(logsp), rtol=1e-9)
    # fast_fftfilt = circular convolution / fft_size (both inputs are scaled by 1 / fft_size)
    full = np.convolve(x, h)
    assert np.allclose(y1[:1000], full / 1024.0, atol=1e-12)


def test_aperiodicity_codec_bit_exact(libs):
    rng = np.random.default_rng(7)
    for fs, n in [(48000, 2048), (16000, 1024), (44100, 2048)]:
        nb = libs[1].GetNumberOfAperiodicities(fs)
        assert libs[0].GetNumberOfAperiodicities(fs) == nb
        F = 12
        ap = np.ascontiguousarray(rng.uniform(0.001, 0.999, (F, n // 2 + 1)))
        ap[3] = 1.0 - 1e-12                                  # an unvoiced frame

        def run(l):
            coded = np.zeros((F, nb)); back = np.zeros((F, n // 2 + 1))
            l.CodeAperiodicity(rows(ap), F, fs, n, nb, rows(coded))
            # what a linked caller gets is the DEFINITION's order (fs, number_of_aperiodicities, fft_size)
            l.DecodeAperiodicity(rows(coded), F, fs, nb, n, rows(back))
            return coded, back
        (c1, b1), (c2, b2) = both(libs, run)
        assert np.array_equal(c1, c2) and np.array_equal(b1, b2)
        assert np.all(b1[3] == 1.0 - 1e-12)

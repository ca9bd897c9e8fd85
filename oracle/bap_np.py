"""TEST INFRASTRUCTURE (oracle): the coded-aperiodicity decode of the reference's synth tool.

W/test/synth.cpp:221-247 turns a coded aperiodicity row (float32 file, `ap_dimension` values) into
aperiodicity bins with SPTK's mgc2sp (W/test/sptkfunctions.cpp:186-275):

    c[0] += 9.210340
    mgc2sp(c, m, ALPHA = 0.55, gamma = 0, x, y, fft_size)    with m = ap_dimension, or ap_dimension - 1 when odd
        = freqt(c, m, c2, fft_size / 2, -ALPHA)               (mgc2mgc with a2 = g1 = g2 = 0; the gnorm / gc2gc /
                                                               ignorm steps are the identity for gamma = 0)
          x = Re FFT_fft_size(c2 zero-padded)                 (c2sp)
    ap[j] = exp(x[j]) / 1e4   for j < m

and leaves the bins j >= m of the row uninitialised; with an even ap_dimension it also reads c[m], one
value past the ones it loaded.  This module restates the DEFINED part: m + 1 coefficients c[0 .. m] in,
x[0 .. fft_size / 2] out; the product fills the whole row with exp(x[j]) / 1e4 and takes c[m] = 0 when
the file does not hold it.  tests/test_bap_decode.py pins this restatement to the compiled reference
routine (oracle/_ref/libsptk_ref.so, built from the reference's own sptkfunctions.cpp by oracle/Makefile).
Only tests/ may import this module."""
import ctypes as C
import os

import numpy as np

ALPHA = 0.55
C0_SHIFT = 9.210340


def freqt(c1, m2, a):
    """W/test/sptkfunctions.cpp freqt: frequency transformation of a cepstrum, order len(c1) - 1 -> m2."""
    c1 = np.asarray(c1, np.float64)
    m1 = len(c1) - 1
    b = 1.0 - a * a
    g = np.zeros(m2 + 1)
    d = np.zeros(m2 + 1)
    for i in range(-m1, 1):
        d[0] = g[0]
        g[0] = c1[-i] + a * d[0]
        if m2 >= 1:
            d[1] = g[1]
            g[1] = b * d[0] + a * d[1]
        for j in range(2, m2 + 1):
            d[j] = g[j]
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
    return g


def decode_row(coded, fft_size):
    """coded: the m + 1 coefficients mgc2sp reads (c0 still carries the tool's -9.210340).  -> x[0..fft_size/2]
    (log spectrum) and ap = exp(x) / 1e4."""
    c = np.asarray(coded, np.float64).copy()
    c[0] += C0_SHIFT
    c2 = freqt(c, fft_size // 2, -ALPHA)
    c2[0] = np.log(np.exp(c2[0]))                       # gnorm / ignorm with gamma = 0
    x = np.fft.rfft(c2, fft_size).real
    return x, np.exp(x) / 1e4


def decode_matrix(m, fft_size):
    """The decode is linear up to the exponential: x = B c.  B[j][i], j <= fft_size / 2, i <= m."""
    B = np.zeros((fft_size // 2 + 1, m + 1))
    for i in range(m + 1):
        e = np.zeros(m + 1)
        e[i] = 1.0
        B[:, i] = np.fft.rfft(freqt(e, fft_size // 2, -ALPHA), fft_size).real
    return B


def reference_mgc2sp(coded, fft_size):
    """The compiled reference routine on the same coefficients -> x[0 .. fft_size)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libsptk_ref.so")
    lib = C.CDLL(path)
    fn = getattr(lib, "_Z6mgc2spPdiddS_S_i")
    dp = C.POINTER(C.c_double)
    fn.argtypes = [dp, C.c_int, C.c_double, C.c_double, dp, dp, C.c_int]
    fn.restype = None
    c = np.ascontiguousarray(coded, np.float64).copy()
    c[0] += C0_SHIFT
    x = np.zeros(fft_size)
    y = np.zeros(fft_size)
    fn(c.ctypes.data_as(dp), len(c) - 1, ALPHA, 0.0, x.ctypes.data_as(dp), y.ctypes.data_as(dp), fft_size)
    return x

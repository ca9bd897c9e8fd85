"""Precision budget of D4C's transforms (CPU study; test infrastructure: it runs the numpy oracle).

DESIGN.md section 4 lists which transform of each stage runs in FP64 and which in FP32 on the
GPU.  This module re-runs the oracle's D4C (oracle/world_np.py, W/src/d4c.cpp:337-397) with
selected transforms emulated in single precision (numpy >= 2 computes float32 / complex64 FFTs in
single precision) and reports the aperiodicity error against the all-double run, so that a
precision change in a kernel can be judged BEFORE it is written:

    python tests/precision_study.py [--quick]        # 12 signals, ~45 s
    python tests/precision_study.py --kernels        # the FP32 transforms today's kernels use, on the hard inputs

Variants of the centroid pair (GetCentroid :90-119; X = FFT(v), Xt = FFT((n + 1) v)):
  f64        : the reference arithmetic
  c32        : both transforms in FP32, weights (n + 1) as in the reference
  c32_centred: FP32, weights (n - c0) with c0 = the window centre; the reference's value is
               recovered as Re(X conj Xt') + (c0 + 1) |X|^2 (the large common part of the two
               terms no longer passes through the FP32 transform of the weighted sequence)
  p32        : the power-spectrum transform (GetSmoothedPowerSpectrum :148-164) in FP32
  b32        : the band transforms (GetCoarseAperiodicity :192-223) in FP32 (what the kernel does)
  lt32       : LoveTrain's transform (:225-250) in FP32
and the power-spectrum transform of CheapTrick (W/src/cheaptrick.cpp:64-82) in FP32 (log spectral
distance in dB, tolerance 0.01).  tests/test_precision_budget.py pins the conclusions the kernels
rely on; profiles/README.md ("Precision budget") holds the table of the last run.
"""
import argparse
import contextlib
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import world_np as W   # noqa: E402


def _rfft(v, n, single):
    if single:
        return np.fft.rfft(v.astype(np.float32), n).astype(np.complex128)
    return np.fft.rfft(v, n)


def d4c_variant(x, fs, t, f0, fft_size, centroid="f64", power32=False, band32=False, lt32=False, threshold=0.0,
                return_ap0=False):
    """oracle.world_np.d4c with selectable transform precision (same structure, same randn stream)."""
    half_out = fft_size // 2
    ap = np.full((len(f0), half_out + 1), 1.0 - W.kMySafeGuardMinimum)
    Nd = W._pow2_above(4.0 * fs / W.kFloorF0D4C + 1)
    nb = int(min(W.kUpperLimit, fs / 2.0 - W.kFrequencyInterval) / W.kFrequencyInterval)
    wl = int(W.kFrequencyInterval * Nd / fs) * 2 + 1
    win = W.nuttall(wl)
    Nl = W._pow2_above(3.0 * fs / 40.0 + 1)
    b0, b1, b2 = (int(math.ceil(v * Nl / fs)) for v in (100.0, 4000.0, 7900.0))
    voiced = np.nonzero(f0 != 0.0)[0]
    n_lt = sum(2 * W.matlab_round(1.5 * fs / max(f0[i], 40.0)) + 1 for i in voiced)
    n_main = sum(3 * (2 * W.matlab_round(2.0 * fs / max(f0[i], W.kFloorF0D4C)) + 1) for i in voiced)
    rn = W.randn_stream(n_lt + n_main)
    pos = 0
    ap0 = np.zeros(len(f0))
    for i in voiced:
        cf = max(f0[i], 40.0)
        w = W._d4c_window(x, fs, cf, t[i], "blackman", 3.0, rn[pos:])
        pos += len(w)
        P = np.abs(_rfft(w, Nl, lt32)) ** 2
        P[:b0 + 1] = 0.0
        c = np.cumsum(P)
        ap0[i] = c[b1] / c[b2]
    coarse_axis = np.concatenate([np.arange(nb + 1) * W.kFrequencyInterval, [fs / 2.0]])
    axis = np.arange(half_out + 1) * fs / fft_size
    for i in voiced:
        if ap0[i] <= threshold:
            continue
        cf = max(W.kFloorF0D4C, f0[i])
        hw = W.matlab_round(2.0 * fs / cf)
        Wn = 2 * hw + 1
        cen = np.zeros(Nd // 2 + 1)
        for side in (-1.0, 1.0):
            w = W._d4c_window(x, fs, cf, t[i] + side * 0.25 / cf, "blackman", 4.0, rn[pos:])
            pos += Wn
            w = w / math.sqrt(np.sum(w * w))
            n = np.arange(Wn, dtype=np.float64)
            if centroid == "f64":
                X, Xt = np.fft.rfft(w, Nd), np.fft.rfft(w * (n + 1.0), Nd)
                cen += X.real * Xt.real + X.imag * Xt.imag
            elif centroid == "c32":
                X, Xt = _rfft(w, Nd, True), _rfft(w * (n + 1.0), Nd, True)
                cen += X.real * Xt.real + X.imag * Xt.imag
            elif centroid == "c32_centred":
                X, Xt = _rfft(w, Nd, True), _rfft(w * (n - hw), Nd, True)
                cen += X.real * Xt.real + X.imag * Xt.imag + (hw + 1.0) * (X.real ** 2 + X.imag ** 2)
            elif centroid == "x64_t32_centred":       # X in FP64, only the weighted sequence in FP32
                X, Xt = np.fft.rfft(w, Nd), _rfft(w * (n - hw), Nd, True)
                cen += X.real * Xt.real + X.imag * Xt.imag + (hw + 1.0) * (X.real ** 2 + X.imag ** 2)
            else:
                raise ValueError(centroid)
        cen = W.dc_correction(cen, cf, fs, Nd)
        w = W._d4c_window(x, fs, cf, t[i], "hanning", 4.0, rn[pos:])
        pos += Wn
        pw = np.abs(_rfft(w, Nd, power32)) ** 2
        pw = W.linear_smoothing(W.dc_correction(pw, cf, fs, Nd), cf, fs, Nd)
        tg = cen / pw
        tg = W.linear_smoothing(tg, cf / 2.0, fs, Nd)
        tg = tg - W.linear_smoothing(tg, cf, fs, Nd)
        boundary = W.matlab_round(Nd * 8.0 / wl)
        coarse = np.zeros(nb + 2)
        coarse[0], coarse[-1] = -60.0, -W.kMySafeGuardMinimum
        for b in range(nb):
            center = int(W.kFrequencyInterval * (b + 1) * Nd / fs)
            seg = tg[center - wl // 2:center - wl // 2 + wl] * win
            P = np.sort(np.abs(_rfft(seg, Nd, band32)) ** 2)
            c = np.cumsum(P)
            coarse[b + 1] = min(0.0, 10 * math.log10(c[Nd // 2 - boundary - 1] / c[Nd // 2]) + (cf - 100.0) / 50.0)
        ap[i] = 10.0 ** (W.interp1(coarse_axis, coarse, axis) / 20.0)
    return (ap, ap0) if return_ap0 else ap


def cheaptrick_variant(x, fs, t, f0, power32=False):
    """oracle.world_np.cheaptrick with the power-spectrum transform (GetPowerSpectrum :64-82, the one
    whose input is shorter than the transform) optionally in single precision."""
    real = np.fft.rfft

    def patched(a, n=None, *args, **kw):
        if power32 and n is not None and len(a) != n:
            return real(np.asarray(a).astype(np.float32), n).astype(np.complex128)
        return real(a, n, *args, **kw)
    np.fft.rfft = patched
    try:
        return W.cheaptrick(x, fs, t, f0)
    finally:
        np.fft.rfft = real


@contextlib.contextmanager
def single_precision_ffts(select=lambda kind, a, n: True):
    """numpy.fft.{rfft, irfft, fft} compute in single precision for the calls `select` accepts."""
    orig = {k: getattr(np.fft, k) for k in ("rfft", "irfft", "fft")}

    def wrap(kind):
        f = orig[kind]

        def g(a, n=None, *args, **kw):
            a = np.asarray(a)
            if not select(kind, a, n):
                return f(a, n, *args, **kw)
            r = f(a.astype(np.complex64 if np.iscomplexobj(a) else np.float32), n, *args, **kw)
            return r.astype(np.complex128 if np.iscomplexobj(r) else np.float64)
        return g
    for k in orig:
        setattr(np.fft, k, wrap(k))
    try:
        yield
    finally:
        for k, f in orig.items():
            setattr(np.fft, k, f)


def kernels_fp32_choices(x, fs, t, f0):
    """The FP32 transforms the kernels run TODAY, emulated one stage at a time
    on the oracle: StoneMask's spectra, CheapTrick's two liftering transforms, Synthesis' four
    transforms.  -> the north_star metrics of each against the all-double oracle."""
    from oracle import metrics as M
    fft_size = W.cheaptrick_fft_size(fs)
    out = {}
    f0_raw = W.dio(x, fs)[1] if f0 is None else f0
    ref = W.stonemask(x, fs, t, f0_raw)
    with single_precision_ffts():
        new = W.stonemask(x, fs, t, f0_raw)
    out["stonemask: V/UV agreement"] = M.vuv_agreement(ref, new)
    out["stonemask: F0 relative error (tol 1e-4)"] = M.f0_rel_error(ref, new)
    sp = W.cheaptrick(x, fs, t, ref)
    with single_precision_ffts(lambda kind, a, n: kind == "irfft" or (kind == "rfft" and (n is None or len(a) == n))):
        sp32 = W.cheaptrick(x, fs, t, ref)
    out["cheaptrick lifter: LSD dB (tol 0.01)"] = M.lsd_db(sp, sp32)[1]
    ap = W.d4c(x, fs, t, ref, fft_size, threshold=0.0)
    y = W.synthesis(ref, sp, ap, fft_size, 5.0, fs)
    with single_precision_ffts():
        y32 = W.synthesis(ref, sp, ap, fft_size, 5.0, fs)
    out["synthesis: SNR dB (tol >= 60)"] = M.snr_db(y, y32)
    return out


def _contour(x, fs):
    t, f0 = W.dio(x, fs)
    return t, W.stonemask(x, fs, t, f0)


def cases(quick=False):
    """(name, x, fs, t, f0): bench-corpus utterances (hts-train-world_b200/signals.py), the real-speech
    fixtures of the reference (tests/golden, written by the compiled reference) and the hard inputs:
    a recording band-limited far below fs/2 (its upper bands hold quantisation noise only), a DC
    offset, a very quiet recording and voiced frames over digital silence."""
    import hts_train_world_b200.signals as S

    def synth(u, fs, seconds):
        pcm, _ = S.make_utterance(u, fs, duration=seconds)
        return pcm.numpy().astype(np.float64) / 32768.0

    def golden(name):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
        return g["pcm"].astype(np.float64) / 32768.0, int(g["fs"]), g["t"], g["f0"]

    out = []
    x = synth(7, 48000, 1.0 if quick else 2.0)
    out.append(("synthetic 48 kHz u7", x, 48000) + _contour(x, 48000))
    x, fs, t, f0 = golden("vaiueo2d")
    out.append(("vaiueo2d 22.05 kHz (real speech)", x, fs, t, f0))
    if quick:
        return out
    x, fs, t, f0 = golden("arctic_a0001")
    out.append(("arctic_a0001 16 kHz (real speech)", x, fs, t, f0))
    from scipy.signal import resample_poly
    x48 = np.round(np.clip(resample_poly(x, 3, 1), -1, 1) * 32767.0) / 32768.0
    out.append(("arctic_a0001 resampled to 48 kHz (empty above 8 kHz)", x48, 48000) + _contour(x48, 48000))
    x = synth(3, 48000, 2.0)
    out.append(("synthetic 48 kHz u3 + DC offset 0.3", x * 0.5 + 0.3, 48000) + _contour(x * 0.5 + 0.3, 48000))
    xq = np.round(x * 32768.0 / 256.0) / 32768.0
    out.append(("synthetic 48 kHz u3 at -48 dB (7-bit)", xq, 48000) + _contour(xq, 48000))
    t = np.arange(201) * 0.005
    out.append(("digital silence, f0 = 150 Hz on every frame", np.zeros(48000), 48000, t, np.full(201, 150.0)))
    for u in (0, 1, 2, 4, 5):          # bench-corpus utterances with the lowest / highest base F0
        x = synth(u, 48000, 1.5)
        out.append(("synthetic 48 kHz u%d" % u, x, 48000) + _contour(x, 48000))
    return out


VARIANTS = [
    ("b32 (kernel today)", dict(band32=True)),
    ("lt32", dict(lt32=True)),
    ("p32", dict(power32=True)),
    ("c32", dict(centroid="c32")),
    ("c32_centred", dict(centroid="c32_centred")),
    ("x64_t32_centred", dict(centroid="x64_t32_centred")),
    ("all FP32: c32 + p32 + b32 + lt32", dict(centroid="c32", power32=True, band32=True, lt32=True)),
]


def lsd_db(a, b):
    d = 10.0 * np.log10(a / b)
    return float(np.max(np.sqrt(np.mean(d * d, axis=1))))


def run(quick=False, variants=VARIANTS):
    """-> {case: {variant: max |ap - ap_f64|, 'ap0 lt32': max |ap0 - ap0_f64|, 'cheaptrick p32': LSD dB}}"""
    res = {}
    for name, x, fs, t, f0 in cases(quick):
        fft_size = W.cheaptrick_fft_size(fs)
        ref, ap0 = d4c_variant(x, fs, t, f0, fft_size, return_ap0=True)
        r = {}
        for vname, kw in variants:
            r[vname] = float(np.max(np.abs(d4c_variant(x, fs, t, f0, fft_size, **kw) - ref)))
        r["ap0 lt32"] = float(np.max(np.abs(d4c_variant(x, fs, t, f0, fft_size, lt32=True, return_ap0=True)[1] - ap0)))
        r["cheaptrick p32 (LSD dB)"] = lsd_db(cheaptrick_variant(x, fs, t, f0, True), cheaptrick_variant(x, fs, t, f0, False))
        r["frames"] = "%d voiced of %d" % (int(np.count_nonzero(f0)), len(f0))
        res[name] = r
    return res


if __name__ == "__main__":
    a = argparse.ArgumentParser()
    a.add_argument("--quick", action="store_true")
    a.add_argument("--kernels", action="store_true", help="the FP32 choices of today's kernels on the hard inputs")
    args = a.parse_args()
    if args.kernels:
        for name, x, fs, t, f0 in cases(False)[:4]:
            print(name)
            for k, v in kernels_fp32_choices(x, fs, t, f0).items():
                print("    %-44s %.4g" % (k, v))
        sys.exit(0)
    for case, r in run(args.quick).items():
        print("%s  [%s]" % (case, r.pop("frames")))
        for k, v in r.items():
            print("    %-36s %.3e" % (k, v))

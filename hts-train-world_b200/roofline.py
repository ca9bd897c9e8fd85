"""Algorithmic FLOP and byte counts per unit of work for every stage of the WORLD path
(DESIGN.md section 5 states the same formulas).  "Algorithmic" = the reference
algorithm's arithmetic, not what our kernels happen to execute: a real FFT of n points counts
2.5 n log2 n, a complex one 5 n log2 n, every elementwise add / multiply / transcendental
counts 1, comparisons count 0.  Bytes are compulsory HBM traffic at the API's precision: every
input read once, every output written once.

Every entry also says in which precision the kernel EXECUTES those operations (`flops64` +
`flops32` = `flops`; DESIGN.md section 4 lists the choice per transform), so that a kernel is
measured against the peak of the pipe it uses: the time bound of a kernel is
flops64 / peak_fp64 + flops32 / peak_fp32 (`roof_seconds`), never all of it over the FP64 peak.
"""
import math

import numpy as np

K_FLOOR_F0_D4C = 47.0


def rfft_flops(n):
    return 2.5 * n * math.log2(n)


def cfft_flops(n):
    return 5.0 * n * math.log2(n)


def _pow2_above(v):
    return int(2 ** (1 + int(math.log(v) / math.log(2.0))))


def cheaptrick_fft_size(fs, f0_floor=71.0):          # W/src/cheaptrick.cpp:191-194
    return _pow2_above(3.0 * fs / f0_floor + 1)


def d4c_fft_size(fs):                                # W/src/d4c.cpp:344-346
    return _pow2_above(4.0 * fs / K_FLOOR_F0_D4C + 1)


def lovetrain_fft_size(fs):                          # W/src/d4c.cpp:261-262
    return _pow2_above(3.0 * fs / 40.0 + 1)


def d4c_bands(fs):                                   # W/src/d4c.cpp:351-353
    return int(min(15000.0, fs / 2.0 - 3000.0) / 3000.0)


def roof_seconds(cnt, peak_fp64_tflops, peak_fp32_tflops, hbm_gbs):
    """Lower bound of a kernel's time: the slower of its arithmetic on the pipes it uses and its
    compulsory bytes at the HBM bandwidth.  -> (seconds, "fp64" | "fp32" | "fp64+fp32" | "hbm")."""
    f32 = float(cnt.get("flops32", 0.0))
    f64 = float(cnt["flops"]) - f32
    t_flop = f64 / (peak_fp64_tflops * 1e12) + f32 / (peak_fp32_tflops * 1e12)
    t_byte = float(cnt["bytes"]) / (hbm_gbs * 1e9)
    if t_byte > t_flop:
        return t_byte, "hbm"
    share32 = f32 / max(1.0, f32 + f64)
    return t_flop, ("fp64" if share32 < 0.1 else "fp32" if share32 > 0.9 else "fp64+fp32")


def stage_counts(fs, f0, n_samples, n_pulses_voiced=None, n_pulses_unvoiced=None, fft_size=None,
                 n_utt=1, dio_fft_sizes=None, lovetrain_fp32=False, ndim_mgc=50, ndim_bap=24):
    """f0: refined F0 of every frame of the batch (numpy).  Returns
    {stage: dict(units, flops, bytes)} with the totals over the batch."""
    f0 = np.asarray(f0, np.float64)
    F = len(f0)
    N = fft_size or cheaptrick_fft_size(fs)
    H = N // 2 + 1
    voiced = f0 > 0
    fv = f0[voiced]
    out = {}
    # ---- CheapTrick: every frame; 3 real FFTs of N; window W = 2 round(1.5 fs / f0') + 1
    f0c = np.where(f0 <= 3.0 * fs / (N - 3.0), 500.0, f0)
    W = 2 * np.round(1.5 * fs / f0c) + 1
    ct = F * (3 * rfft_flops(N) + 36.0 * H) + 12.0 * W.sum()
    # FP32: the two liftering transforms (log spectrum <-> cepstrum) and the lifter itself (6 per bin)
    ct32 = F * (2 * rfft_flops(N) + 6.0 * H)
    out["cheaptrick"] = dict(units=F, flops=ct, flops32=ct32, bytes=8.0 * n_samples + F * (16.0 + 8.0 * H))
    # ---- D4C: voiced frames only
    Nd, Nl, nb = d4c_fft_size(fs), lovetrain_fft_size(fs), d4c_bands(fs)
    Hd = Nd // 2 + 1
    W3 = 2 * np.round(1.5 * fs / np.maximum(fv, 40.0)) + 1
    W4 = 2 * np.round(2.0 * fs / np.maximum(fv, K_FLOOR_F0_D4C)) + 1
    wl = int(3000.0 * Nd / fs) * 2 + 1
    lt = len(fv) * (rfft_flops(Nl) + 3.0 * (Nl // 2 + 1)) + 14.0 * W3.sum()
    main = len(fv) * (10 * rfft_flops(Nd) + 50.0 * Hd + nb * (4.0 * Hd + wl)) + 48.0 * W4.sum()
    # FP32: the nb band transforms (one real FFT each), their powers and the selection; LoveTrain's
    # transform when lovetrain_fp32 (the window, the mean removal and the band sums stay FP64)
    main32 = len(fv) * (nb * rfft_flops(Nd) + nb * (4.0 * Hd + wl))
    lt32 = len(fv) * rfft_flops(Nl) if lovetrain_fp32 else 0.0
    out["d4c_lovetrain"] = dict(units=int(len(fv)), flops=lt, flops32=lt32, bytes=8.0 * n_samples + 16.0 * F)
    out["d4c_main"] = dict(units=int(len(fv)), flops=main, flops32=main32, bytes=8.0 * n_samples + F * (16.0 + 8.0 * H))
    # the two kernels of the split layout: the FP64 group-delay kernel writes nb windowed slices of
    # wl floats per voiced frame, the FP32 tail reads them and writes the aperiodicity row
    out["d4c_gd"] = dict(units=int(len(fv)), flops=main - main32, flops32=0.0,
                         bytes=8.0 * n_samples + 16.0 * F + 4.0 * nb * wl * len(fv))
    out["d4c_tail"] = dict(units=int(len(fv)), flops=main32, flops32=main32,
                           bytes=4.0 * nb * wl * len(fv) + F * (16.0 + 8.0 * H))
    out["d4c"] = dict(units=int(len(fv)), flops=lt + main, flops32=lt32 + main32, bytes=8.0 * n_samples + F * (16.0 + 8.0 * H))
    # ---- StoneMask: voiced frames; 2 real FFTs of 2^(2 + floor(log2(2 hwl + 1)))
    sm = fv[(fv > 40.0) & (fv <= fs / 12.0)]
    hwl = np.floor(1.5 * fs / sm + 1.0)
    nfft = 2.0 ** (2 + np.floor(np.log2(2 * hwl + 1)))
    out["stonemask"] = dict(units=int(len(sm)),
                            flops=float((2 * 2.5 * nfft * np.log2(nfft) + 20.0 * (2 * hwl + 1)).sum()),
                            flops32=float((2 * 2.5 * nfft * np.log2(nfft)).sum()),      # the transform; windows / FixF0 are FP64
                            bytes=8.0 * n_samples + 24.0 * F)
    # ---- Dio: 16 real FFTs of fft_size per utterance + 7 bands x (complex multiply, 4
    #      zero-crossing passes, 4 interp1 per frame)
    if dio_fft_sizes is None:
        per = max(2, n_samples // max(1, n_utt))
        dio_fft_sizes = [_pow2_above(per + 1 + 4 * int(1 + fs / (71.0 * 2 ** 0.5) / 2.0))] * n_utt
    dio = sum(16 * rfft_flops(n) + 7 * 3.0 * n for n in dio_fft_sizes) + 7 * 16.0 * n_samples + 7 * 60.0 * F
    out["dio"] = dict(units=n_utt, flops=dio, flops32=0.0, bytes=8.0 * n_samples + 16.0 * F)
    # ---- codec tail (W/src/codec.cpp:266-295, twice per frame: mgc from sp, bap from ap): log of every
    #      bin, interp1 onto fft_size/2 mel points (6 per point), one real FFT of fft_size/2, scalings
    Nc = N // 2
    per = H + 6.0 * Nc + rfft_flops(Nc)
    out["codec"] = dict(units=2 * F, flops=2 * F * per + F * (ndim_mgc + ndim_bap + 1.0), flops32=2.0 * F * H,   # the batch path's FP32 logarithm
                        bytes=F * (16.0 * H + 8.0 + 4.0 * (ndim_mgc + ndim_bap + 1)))
    # ---- Synthesis: voiced pulse 3 r2c + 2 c2c + 2 c2r of N + ~60k; unvoiced 2 r2c + c2c + c2r + ~45k
    if n_pulses_voiced is not None:
        pv, pu = float(n_pulses_voiced), float(n_pulses_unvoiced)
        scale = N / 2048.0
        syn = pv * (5 * rfft_flops(N) + 2 * cfft_flops(N) + 60e3 * scale) + \
            pu * (3 * rfft_flops(N) + cfft_flops(N) + 45e3 * scale) + 30.0 * n_samples
        # FP32: the transforms and the spectral arithmetic between them (log / exp / products of the
        # two-channel work items); FP64: the time base (30 per sample), the sp / ap interpolation of
        # every pulse (8 per bin and channel) and the overlap-add (N per pulse)
        syn64 = 30.0 * n_samples + (pv + pu) * (8.0 * H + N)
        out["synthesis"] = dict(units=int(pv + pu), flops=syn, flops32=max(0.0, syn - syn64),
                                bytes=F * (8.0 + 16.0 * H) + 8.0 * n_samples)
    return out

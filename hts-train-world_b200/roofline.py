"""Algorithmic FLOP and byte counts per unit of work for every stage of the WORLD path
(DESIGN.md section 5 states the same formulas).  "Algorithmic" = the reference
algorithm's arithmetic, not what our kernels happen to execute: a real FFT of n points counts
2.5 n log2 n, a complex one 5 n log2 n, every elementwise add / multiply / transcendental
counts 1, comparisons count 0.  Bytes are compulsory HBM traffic at the API's precision: every
input read once, every output written once.
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


def stage_counts(fs, f0, n_samples, n_pulses_voiced=None, n_pulses_unvoiced=None, fft_size=None,
                 n_utt=1, dio_fft_sizes=None):
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
    out["cheaptrick"] = dict(units=F, flops=ct, bytes=8.0 * n_samples + F * (16.0 + 8.0 * H))
    # ---- D4C: voiced frames only
    Nd, Nl, nb = d4c_fft_size(fs), lovetrain_fft_size(fs), d4c_bands(fs)
    Hd = Nd // 2 + 1
    W3 = 2 * np.round(1.5 * fs / np.maximum(fv, 40.0)) + 1
    W4 = 2 * np.round(2.0 * fs / np.maximum(fv, K_FLOOR_F0_D4C)) + 1
    wl = int(3000.0 * Nd / fs) * 2 + 1
    lt = len(fv) * (rfft_flops(Nl) + 3.0 * (Nl // 2 + 1)) + 14.0 * W3.sum()
    main = len(fv) * (10 * rfft_flops(Nd) + 50.0 * Hd + nb * (4.0 * Hd + wl)) + 48.0 * W4.sum()
    out["d4c_lovetrain"] = dict(units=int(len(fv)), flops=lt, bytes=8.0 * n_samples + 16.0 * F)
    out["d4c_main"] = dict(units=int(len(fv)), flops=main, bytes=8.0 * n_samples + F * (16.0 + 8.0 * H))
    out["d4c"] = dict(units=int(len(fv)), flops=lt + main, bytes=8.0 * n_samples + F * (16.0 + 8.0 * H))
    # ---- StoneMask: voiced frames; 2 real FFTs of 2^(2 + floor(log2(2 hwl + 1)))
    sm = fv[(fv > 40.0) & (fv <= fs / 12.0)]
    hwl = np.floor(1.5 * fs / sm + 1.0)
    nfft = 2.0 ** (2 + np.floor(np.log2(2 * hwl + 1)))
    out["stonemask"] = dict(units=int(len(sm)),
                            flops=float((2 * 2.5 * nfft * np.log2(nfft) + 20.0 * (2 * hwl + 1)).sum()),
                            bytes=8.0 * n_samples + 24.0 * F)
    # ---- Dio: 16 real FFTs of fft_size per utterance + 7 bands x (complex multiply, 4
    #      zero-crossing passes, 4 interp1 per frame)
    if dio_fft_sizes is None:
        per = max(2, n_samples // max(1, n_utt))
        dio_fft_sizes = [_pow2_above(per + 1 + 4 * int(1 + fs / (71.0 * 2 ** 0.5) / 2.0))] * n_utt
    dio = sum(16 * rfft_flops(n) + 7 * 3.0 * n for n in dio_fft_sizes) + 7 * 16.0 * n_samples + 7 * 60.0 * F
    out["dio"] = dict(units=n_utt, flops=dio, bytes=8.0 * n_samples + 16.0 * F)
    # ---- Synthesis: voiced pulse 3 r2c + 2 c2c + 2 c2r of N + ~60k; unvoiced 2 r2c + c2c + c2r + ~45k
    if n_pulses_voiced is not None:
        pv, pu = float(n_pulses_voiced), float(n_pulses_unvoiced)
        scale = N / 2048.0
        syn = pv * (5 * rfft_flops(N) + 2 * cfft_flops(N) + 60e3 * scale) + \
            pu * (3 * rfft_flops(N) + cfft_flops(N) + 45e3 * scale) + 30.0 * n_samples
        out["synthesis"] = dict(units=int(pv + pu), flops=syn,
                                bytes=F * (8.0 + 16.0 * H) + 8.0 * n_samples)
    return out

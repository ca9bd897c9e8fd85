"""TEST INFRASTRUCTURE.  numpy restatement of ONE piece of Harvest: the refinement of the overlapped F0 candidates by
instantaneous frequency (W/src/harvest.cpp: OverlapF0Candidates :417-429, GetBaseIndex :434-441, GetMainWindow :446-456,
GetDiffWindow :462-468, GetSpectra :474-504, FixF0 :506-536, GetMeanF0 :541-582, GetRefinedF0 :587-616,
RefineF0Candidates :621-631).  It exists to check the arithmetic of harvest_refine_thread_kernel (Goertzel recurrences
instead of two FFTs per candidate) on the CPU, through tests/emu.  The functions are file-static in the reference, so
this piece cannot be pinned to the compiled library on its own: parity unpinned for it; Harvest AS A WHOLE is pinned to
the compiled reference by the GPU parity tests (tests/test_gpu_parity.py::test_harvest_*).  Never imported by the product.
"""
import math

import numpy as np

kPi = 3.1415926535897932384
kMySafeGuardMinimum = 0.000000000001
kLog2 = 0.69314718055994529


def matlab_round(x):                                    # W/src/matlabfunctions.cpp:212-214
    return int(x + 0.5) if x > 0 else int(x - 0.5)


def overlap_candidates(base, nc):
    """base: [n_fr][>= nc] -> [n_fr][7 nc] (:417-429; slots that no frame fills stay 0)."""
    n_fr = base.shape[0]
    out = np.zeros((n_fr, 7 * nc))
    out[:, :nc] = base[:, :nc]
    for i in range(1, 4):
        out[i:, nc * i:nc * (i + 1)] = base[:n_fr - i, :nc]
        out[:n_fr - i, nc * (i + 3):nc * (i + 4)] = base[i:, :nc]
    return out


def refined_f0(x, fs, position, f0, f0_floor, f0_ceil):
    """GetRefinedF0 (:587-616) for one candidate; x = the decimated signal with its mean removed."""
    if f0 <= 0.0:
        return 0.0, 0.0
    hwl = int(1.5 * fs / f0 + 1.0)
    W = 2 * hwl + 1
    wlen = (2.0 * hwl + 1.0) / fs
    fft_size = int(2.0 ** (2.0 + int(math.log(hwl * 2.0 + 1.0) / kLog2)))
    basic_index = matlab_round((position + (-hwl) / fs) * fs + 0.001)          # :437-438
    idx = basic_index + np.arange(W)
    tmp = (idx - 1.0) / fs - position
    main = 0.42 + 0.5 * np.cos(2.0 * kPi * tmp / wlen) + 0.08 * np.cos(4.0 * kPi * tmp / wlen)   # :451-455
    diff = np.empty(W)
    diff[0] = -main[1] / 2.0
    diff[1:-1] = -(main[2:] - main[:-2]) / 2.0
    diff[-1] = main[-2] / 2.0
    safe = np.clip(idx - 1, 0, len(x) - 1)
    ms = np.fft.rfft(x[safe] * main, fft_size)
    ds = np.fft.rfft(x[safe] * diff, fft_size)
    numerator_i = ms.real * ds.imag - ms.imag * ds.real                        # :558-561
    power = ms.real ** 2 + ms.imag ** 2
    nh = min(int(fs / 2.0 / f0), 6)
    num = den = score = 0.0
    for i in range(nh):                                                        # FixF0 :513-529
        k = matlab_round(f0 * fft_size / fs * (i + 1))
        inst = 0.0 if power[k] == 0.0 else k * fs / fft_size + numerator_i[k] / power[k] * fs / 2.0 / kPi
        amp = math.sqrt(power[k])
        num += amp * inst
        den += amp * (i + 1.0)
        score += abs((inst / (i + 1.0) - f0) / f0)
    r = num / (den + kMySafeGuardMinimum)
    s = 1.0 / (score / nh + kMySafeGuardMinimum)
    if r < f0_floor or r > f0_ceil or s < 2.5:                                 # :607-611
        return 0.0, 0.0
    return r, s


def refine_candidates(y, fs, base, nc, f0_floor, f0_ceil):
    """RefineF0Candidates (:621-631) on the overlapped candidates; frame k sits at k ms.  -> (cand, score) [n_fr][7 nc]."""
    x = np.asarray(y, np.float64) - np.mean(y)
    ov = overlap_candidates(np.asarray(base, np.float64), nc)
    cand, score = np.zeros_like(ov), np.zeros_like(ov)
    for k in range(ov.shape[0]):
        for s in range(ov.shape[1]):
            cand[k, s], score[k, s] = refined_f0(x, fs, k / 1000.0, ov[k, s], f0_floor, f0_ceil)
    return cand, score

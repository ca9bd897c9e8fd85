"""TEST INFRASTRUCTURE ONLY — numpy restatement of the WORLD analysis/synthesis algorithms of
externs/WORLD_v2 (W/ below), written from the reference's sources as an independent second
oracle.  Only tests/ may import this module; the product never does.

Pinned: tests/test_oracle_np.py checks every function here (Dio, StoneMask, CheapTrick, D4C,
Synthesis, codec; Harvest is checked against the compiled reference only) against the golden vectors in
tests/golden/ (produced by the compiled, unmodified reference) — F0 within 1e-6 relative,
spectral envelope within 1e-6 dB, aperiodicity within 1e-7, resynthesis > 100 dB SNR; the
differences are numpy's FFT versus the reference's Ooura FFT.  The strongest oracle remains
the compiled reference itself (oracle/ref.py); this file documents the algorithm in ~400 lines
and lets the tests tell an algorithmic discrepancy from an FFT-rounding one.

The two strictly sequential pieces (xorshift128 randn stream, pulse phase accumulation) are
in oracle/world_port.c, loaded through ctypes.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
kPi = 3.1415926535897932384
kMySafeGuardMinimum = 1e-12
kEps = 2.2204460492503131e-16
kDefaultF0 = 500.0
kFloorF0D4C = 47.0
kFrequencyInterval = 3000.0
kUpperLimit = 15000.0
kM0, kF0mel = 1127.01048, 700.0

_port = None


def port():
    global _port
    if _port is None:
        p = os.path.join(_HERE, "_ref", "libworld_port.so")
        if not os.path.exists(p):
            raise FileNotFoundError(p + " missing: run `make -C oracle _ref/libworld_port.so`")
        _port = C.CDLL(p)
    return _port


def randn_stream(n):
    """W/src/matlabfunctions.cpp:247-277 after randn_reseed()."""
    out = np.zeros(int(n))
    port().port_randn_stream(out.ctypes.data_as(C.POINTER(C.c_double)), C.c_longlong(int(n)))
    return out


def matlab_round(x):                                   # W/src/matlabfunctions.cpp:212-214
    return int(x + 0.5) if x > 0 else int(x - 0.5)


def interp1(x, y, xi):
    """W/src/matlabfunctions.cpp:157-182 (histc :136-155): segment k = clamp(upper_bound, 1, n-1),
    linear extrapolation outside the knots."""
    x, y, xi = np.asarray(x, float), np.asarray(y, float), np.asarray(xi, float)
    k = np.clip(np.searchsorted(x, xi, side="right"), 1, len(x) - 1)
    s = (xi - x[k - 1]) / (x[k] - x[k - 1])
    return y[k - 1] + s * (y[k] - y[k - 1])


def interp1q(x0, dx, y, xi):
    """W/src/matlabfunctions.cpp:220-241: uniform grid, base index by truncation, delta_y[last] = 0."""
    y = np.asarray(y, float)
    r = (np.asarray(xi, float) - x0) / dx
    base = r.astype(np.int64)
    frac = r - base
    dy = np.append(np.diff(y), 0.0)
    return y[base] + dy[base] * frac


def dc_correction(spec, f0, fs, fft_size):             # W/src/common.cpp:56-75
    out = spec.copy()
    upper = 2 + int(f0 * fft_size / fs)
    axis = np.arange(upper) * fs / fft_size
    rep = interp1q(f0 - axis[0], -fs / fft_size, spec[:upper + 1], axis[:upper - 1])
    out[:upper - 1] = spec[:upper - 1] + rep
    return out


def linear_smoothing(spec, width, fs, fft_size):       # W/src/common.cpp:27-46, 77-111
    half = fft_size // 2
    boundary = int(width * fft_size / fs) + 1
    mirror = np.concatenate([spec[boundary:0:-1], spec[:half], spec[half:half - boundary - 1:-1]])
    seg = np.cumsum(mirror * fs / fft_size)
    axis = np.arange(half + 1) / fft_size * fs - width / 2.0
    origin = -(boundary - 0.5) * fs / fft_size
    low = interp1q(origin, fs / fft_size, seg, axis)
    high = interp1q(origin, fs / fft_size, seg, axis + width)
    return (high - low) / width


def nuttall(n):                                        # W/src/common.cpp:113-121
    t = np.arange(n) / (n - 1.0)
    return 0.355768 - 0.487396 * np.cos(2 * kPi * t) + 0.144232 * np.cos(4 * kPi * t) - 0.012604 * np.cos(6 * kPi * t)


def _pow2_above(v):
    return int(2.0 ** (1.0 + int(math.log(v) / math.log(2.0))))


# ---------------------------------------------------------------------------------------------
# CheapTrick (W/src/cheaptrick.cpp)
# ---------------------------------------------------------------------------------------------
def cheaptrick_fft_size(fs, f0_floor=71.0):            # :191-194
    return _pow2_above(3.0 * fs / f0_floor + 1)


def cheaptrick(x, fs, t, f0, q1=-0.15, fft_size=None):
    """:200-228 with per-frame body :159-187.  Returns [frames][fft_size/2+1]."""
    N = fft_size or cheaptrick_fft_size(fs)
    half = N // 2
    f0_floor = 3.0 * fs / (N - 3.0)                    # :196-198
    f0c = np.where(f0 <= f0_floor, kDefaultF0, f0)     # :217
    hwl = np.array([matlab_round(1.5 * fs / v) for v in f0c])
    rn = randn_stream(int((2 * hwl + 1).sum() + len(f0) * (half + 1)))
    out = np.zeros((len(f0), half + 1))
    pos = 0
    for i, (ti, fi, h) in enumerate(zip(t, f0c, hwl)):
        # GetWindowedWaveform :112-142
        base = np.arange(-h, h + 1)
        idx = np.clip(matlab_round(ti * fs + 0.001) + base, 0, len(x) - 1)
        w = 0.5 * np.cos(kPi * (base / 1.5 / fs) * fi) + 0.5
        w = w / math.sqrt(np.sum(w * w))
        wave = x[idx] * w + rn[pos:pos + 2 * h + 1] * kMySafeGuardMinimum
        pos += 2 * h + 1
        wave = wave - w * (wave.sum() / w.sum())
        # GetPowerSpectrum :64-82
        X = np.fft.rfft(wave, N)
        power = dc_correction(X.real ** 2 + X.imag ** 2, fi, fs, N)
        # LinearSmoothing, AddInfinitesimalNoise :147-151
        sm = linear_smoothing(power, fi * 2.0 / 3.0, fs, N)
        sm = sm + np.abs(rn[pos:pos + half + 1]) * kEps
        pos += half + 1
        # SmoothingWithRecovery :22-57
        q = np.arange(half + 1) / fs
        lifter = np.ones(half + 1)
        lifter[1:] = np.sin(kPi * fi * q[1:]) / (kPi * fi * q[1:])
        comp = (1.0 - 2.0 * q1) + 2.0 * q1 * np.cos(2.0 * kPi * q * fi)
        logs = np.log(sm)
        cep = np.fft.rfft(np.concatenate([logs, logs[half - 1:0:-1]])).real
        env = np.fft.irfft(cep * lifter * comp, N)[:half + 1]      # irfft includes the 1/N of :49
        out[i] = np.exp(env)
    return out


# ---------------------------------------------------------------------------------------------
# Dio (W/src/dio.cpp), speed = 1 (the tool's setting; decimation is checked against the compiled
# reference only)
# ---------------------------------------------------------------------------------------------
kMaximumValue = 100000.0
kLog2 = 0.69314718055994529


def _low_cut_filter(N, fft_size):
    """DesignLowCutFilter :40-53, including its in-place shifts (the second one reads past N into
    the zeroed area)."""
    f = np.zeros(fft_size)
    i = np.arange(1, N + 1)
    f[:N] = 0.5 - 0.5 * np.cos(i * 2.0 * kPi / (N + 1))
    f[:N] = -f[:N] / f[:N].sum()
    h = (N - 1) // 2
    f[fft_size - h:fft_size] = f[:h]
    f[:N] = f[h:h + N].copy()          # every read index is above its write index: same as the loop
    f[0] += 1.0
    return f


def _zero_crossing_engine(sig, n, fs):
    """ZeroCrossingEngine :357-393 on sig[0..n): (interval_locations, intervals), empty if < 2 edges."""
    s = sig[:n]
    edges = np.nonzero((0.0 < s[:-1]) & (s[1:] <= 0.0))[0] + 1
    if len(edges) < 2:
        return np.zeros(0), np.zeros(0)
    fine = edges - s[edges - 1] / (s[edges] - s[edges - 1])
    return (fine[:-1] + fine[1:]) / 2.0 / fs, fs / (fine[1:] - fine[:-1])


def _dio_band(Y, fft_size, y_length, fs, boundary_f0, f0_floor, f0_ceil, t):
    """GetF0CandidateFromRawEvent :523-541 -> (candidate, score) of one band."""
    hal = matlab_round(fs / boundary_f0 / 2.0)
    lpf = np.zeros(fft_size)                                              # GetFilteredSignal :296-343
    lpf[:4 * hal] = nuttall(4 * hal)
    conv = np.fft.irfft(Y * np.fft.rfft(lpf), fft_size) * fft_size       # c2r is unnormalised
    sig = conv[2 * hal:2 * hal + y_length].copy()
    sets = []                                                             # GetFourZeroCrossingIntervals :402-435
    sets.append(_zero_crossing_engine(sig, y_length, fs))
    sig = -sig
    sets.append(_zero_crossing_engine(sig, y_length, fs))
    sig[:y_length - 1] = sig[:y_length - 1] - sig[1:y_length]
    sets.append(_zero_crossing_engine(sig, y_length - 1, fs))
    sig[:y_length - 1] = -sig[:y_length - 1]
    sets.append(_zero_crossing_engine(sig, y_length - 1, fs))
    F = len(t)
    if any(len(v) - 2 <= 0 for _, v in sets):                             # CheckEvent :484-493
        return np.zeros(F), np.full(F, kMaximumValue)
    ip = np.array([interp1(loc, val, t) for loc, val in sets])            # :498-512
    cand = (ip[0] + ip[1] + ip[2] + ip[3]) / 4.0                          # GetF0CandidateContourSub :441-465
    score = np.sqrt(((ip - cand) ** 2).sum(axis=0) / 3.0)
    bad = (cand > boundary_f0) | (cand < boundary_f0 / 2.0) | (cand > f0_ceil) | (cand < f0_floor)
    cand[bad] = 0.0
    score[bad] = kMaximumValue
    return cand, score


def _select_best_f0(cur, past, cands, j, allowed):                       # SelectBestF0 :190-209
    ref = (cur * 3.0 - past) / 2.0
    err = np.abs(ref - cands[:, j])
    best = cands[int(np.argmin(err)), j]                                   # first minimum wins
    if abs(1.0 - best / ref) > allowed:
        return 0.0
    return best


def dio(x, fs, frame_period=5.0, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0, allowed_range=0.1):
    """Dio :642-647 / DioGeneralBody :578-634 with speed 1 -> (temporal_positions, f0)."""
    x = np.asarray(x, float)
    x_length = len(x)
    nb = 1 + int(math.log(f0_ceil / f0_floor) / kLog2 * channels_in_octave)           # :582-586
    boundary = [f0_floor * 2.0 ** ((i + 1) / channels_in_octave) for i in range(nb)]
    y_length = 1 + x_length                                                            # :590 (speed 1)
    fft_size = _pow2_above(y_length + 4 * int(1.0 + fs / boundary[0] / 2.0))          # :592-593
    y = np.zeros(fft_size)                                                             # GetSpectrumForEstimation :60-106
    y[:x_length] = x
    y[:y_length] -= y[:y_length].sum() / y_length
    Y = np.fft.rfft(y) * np.fft.rfft(_low_cut_filter(matlab_round(fs / 50.0) * 2 + 1, fft_size))
    F = int(1000.0 * x_length / fs / frame_period) + 1                                 # GetSamplesForDIO :638-640
    t = np.arange(F) * frame_period / 1000.0
    cands, scores = np.zeros((nb, F)), np.zeros((nb, F))
    for b in range(nb):                                                                # GetF0CandidatesAndScores :549-572
        c, sc = _dio_band(Y, fft_size, y_length, float(fs), boundary[b], f0_floor, f0_ceil, t)
        cands[b], scores[b] = c, sc / (c + kMySafeGuardMinimum)
    best = cands[np.argmin(scores, axis=0), np.arange(F)]                              # GetBestF0Contour :112-126 (first minimum)
    # FixF0Contour :259-289
    vrm = int(0.5 + 1000.0 / frame_period / f0_floor) * 2 + 1
    f0 = np.zeros(F)
    if F <= vrm:
        return t, f0                                 # the reference returns without writing f0 at all
    base = np.zeros(F)                                                                 # FixStep1 :132-150
    base[vrm:F - vrm] = best[vrm:F - vrm]
    s1 = np.zeros(F)
    i = np.arange(vrm, F)
    ok = np.abs((base[i] - base[i - 1]) / (kMySafeGuardMinimum + base[i])) < allowed_range
    s1[i] = np.where(ok, base[i], 0.0)
    s2 = s1.copy()                                                                     # FixStep2 :156-169
    center = (vrm - 1) // 2
    for i in range(center, F - center):
        if (s1[i - center:i + center + 1] == 0).any():
            s2[i] = 0.0
    neg = [i - 1 for i in range(1, F) if s2[i] == 0 and s2[i - 1] != 0]                # GetNumberOfVoicedSections :174-184
    pos = [i for i in range(1, F) if s2[i - 1] == 0 and s2[i] != 0]
    s3 = s2.copy()                                                                     # FixStep3 :215-231
    for k, start in enumerate(neg):
        limit = F - 1 if k == len(neg) - 1 else neg[k + 1]
        for j in range(start, limit):
            s3[j + 1] = _select_best_f0(s3[j], s3[j - 1], cands, j + 1, allowed_range)
            if s3[j + 1] == 0:
                break
    s4 = s3.copy()                                                                     # FixStep4 :237-253
    for k in range(len(pos) - 1, -1, -1):
        limit = 1 if k == 0 else pos[k - 1]
        for j in range(pos[k], limit, -1):
            s4[j - 1] = _select_best_f0(s4[j], s4[j + 1], cands, j - 1, allowed_range)
            if s4[j - 1] == 0:
                break
    return t, s4


# ---------------------------------------------------------------------------------------------
# D4C (W/src/d4c.cpp)
# ---------------------------------------------------------------------------------------------
def _d4c_window(x, fs, f0, position, kind, ratio, rn):
    """GetWindowedWaveform :52-84 (+ SetParameters :21-46).  Consumes 2*hwl+1 variates."""
    h = matlab_round(ratio * fs / f0 / 2.0)
    base = np.arange(-h, h + 1)
    idx = np.clip(matlab_round(position * fs + 0.001) + base, 0, len(x) - 1)
    p = 2.0 * base / ratio / fs
    if kind == "hanning":
        w = 0.5 * np.cos(kPi * p * f0) + 0.5
    else:
        w = 0.42 + 0.5 * np.cos(kPi * p * f0) + 0.08 * np.cos(kPi * p * f0 * 2)
    wave = x[idx] * w + rn[:2 * h + 1] * kMySafeGuardMinimum
    return wave - w * (wave.sum() / w.sum())


def d4c(x, fs, t, f0, fft_size, threshold=0.85):
    """:337-397.  Returns [frames][fft_size/2+1]."""
    half_out = fft_size // 2
    ap = np.full((len(f0), half_out + 1), 1.0 - kMySafeGuardMinimum)
    Nd = _pow2_above(4.0 * fs / kFloorF0D4C + 1)       # :344-346
    nb = int(min(kUpperLimit, fs / 2.0 - kFrequencyInterval) / kFrequencyInterval)   # :351-353
    wl = int(kFrequencyInterval * Nd / fs) * 2 + 1     # :356
    win = nuttall(wl)
    # LoveTrain :225-282 (its randn draws come before the main loop's)
    Nl = _pow2_above(3.0 * fs / 40.0 + 1)
    b0, b1, b2 = (int(math.ceil(v * Nl / fs)) for v in (100.0, 4000.0, 7900.0))
    voiced = np.nonzero(f0 != 0.0)[0]
    n_lt = sum(2 * matlab_round(1.5 * fs / max(f0[i], 40.0)) + 1 for i in voiced)
    n_main = sum(3 * (2 * matlab_round(2.0 * fs / max(f0[i], kFloorF0D4C)) + 1) for i in voiced)
    rn = randn_stream(n_lt + n_main)
    pos = 0
    ap0 = np.zeros(len(f0))
    for i in voiced:
        cf = max(f0[i], 40.0)
        w = _d4c_window(x, fs, cf, t[i], "blackman", 3.0, rn[pos:])
        pos += len(w)
        P = np.abs(np.fft.rfft(w, Nl)) ** 2
        P[:b0 + 1] = 0.0
        c = np.cumsum(P)
        ap0[i] = c[b1] / c[b2]
    coarse_axis = np.concatenate([np.arange(nb + 1) * kFrequencyInterval, [fs / 2.0]])   # :362-365
    axis = np.arange(half_out + 1) * fs / fft_size
    for i in voiced:
        if ap0[i] <= threshold:                        # :380
            continue
        cf = max(kFloorF0D4C, f0[i])
        W = 2 * matlab_round(2.0 * fs / cf) + 1
        # GetStaticCentroid :125-142 / GetCentroid :90-119
        cen = np.zeros(Nd // 2 + 1)
        for side in (-1.0, 1.0):
            w = _d4c_window(x, fs, cf, t[i] + side * 0.25 / cf, "blackman", 4.0, rn[pos:])
            pos += W
            w = w / math.sqrt(np.sum(w * w))
            X = np.fft.rfft(w, Nd)
            Xt = np.fft.rfft(w * (np.arange(W) + 1.0), Nd)
            cen += X.real * Xt.real + X.imag * Xt.imag
        cen = dc_correction(cen, cf, fs, Nd)
        # GetSmoothedPowerSpectrum :148-164
        w = _d4c_window(x, fs, cf, t[i], "hanning", 4.0, rn[pos:])
        pos += W
        pw = np.abs(np.fft.rfft(w, Nd)) ** 2
        pw = linear_smoothing(dc_correction(pw, cf, fs, Nd), cf, fs, Nd)
        # GetStaticGroupDelay :170-186
        tg = cen / pw
        tg = linear_smoothing(tg, cf / 2.0, fs, Nd)
        tg = tg - linear_smoothing(tg, cf, fs, Nd)
        # GetCoarseAperiodicity :192-223
        boundary = matlab_round(Nd * 8.0 / wl)
        coarse = np.zeros(nb + 2)
        coarse[0], coarse[-1] = -60.0, -kMySafeGuardMinimum
        for b in range(nb):
            center = int(kFrequencyInterval * (b + 1) * Nd / fs)
            seg = tg[center - wl // 2:center - wl // 2 + wl] * win
            P = np.sort(np.abs(np.fft.rfft(seg, Nd)) ** 2)
            c = np.cumsum(P)
            coarse[b + 1] = min(0.0, 10 * math.log10(c[Nd // 2 - boundary - 1] / c[Nd // 2]) + (cf - 100.0) / 50.0)   # :304-308
        ap[i] = 10.0 ** (interp1(coarse_axis, coarse, axis) / 20.0)      # :325-333
    return ap


# ---------------------------------------------------------------------------------------------
# StoneMask (W/src/stonemask.cpp)
# ---------------------------------------------------------------------------------------------
def stonemask(x, fs, t, f0):
    out = np.zeros(len(f0))
    for i, (ti, fi) in enumerate(zip(t, f0)):
        if fi <= 40.0 or fi > fs / 12.0:               # :186-187
            continue
        h = int(1.5 * fs / fi + 1.0)
        wlen = (2.0 * h + 1.0) / fs
        nfft = int(2.0 ** (2.0 + int(math.log(h * 2.0 + 1.0) / math.log(2.0))))   # :192-193
        base_time = np.arange(-h, h + 1) / fs
        index = np.array([matlab_round((ti + b) * fs) for b in base_time])        # :24-28
        tmp = (index - 1.0) / fs - ti
        w = 0.42 + 0.5 * np.cos(2 * kPi * tmp / wlen) + 0.08 * np.cos(4 * kPi * tmp / wlen)    # :33-43
        dw = np.zeros_like(w)                                                      # :49-55
        dw[0] = -w[1] / 2.0
        dw[1:-1] = -(w[2:] - w[:-2]) / 2.0
        dw[-1] = w[-2] / 2.0
        xs = x[np.clip(index - 1, 0, len(x) - 1)]
        Xm, Xd = np.fft.rfft(xs * w, nfft), np.fft.rfft(xs * dw, nfft)
        power = Xm.real ** 2 + Xm.imag ** 2
        numer = Xm.real * Xd.imag - Xm.imag * Xd.real

        def fix(f_init, nh):                                                       # :96-117
            num = den = 0.0
            for k in range(nh):
                idx = matlab_round(f_init * nfft / fs * (k + 1))
                inst = 0.0 if power[idx] == 0.0 else idx * fs / nfft + numer[idx] / power[idx] * fs / 2.0 / kPi
                amp = math.sqrt(power[idx])
                num += amp * inst
                den += amp * (k + 1)
            return num / (den + kMySafeGuardMinimum)

        tentative = fix(fi, 2)                                                     # :122-131
        mean = 0.0 if (tentative <= 0.0 or tentative > fi * 2) else fix(tentative, 6)
        out[i] = fi if abs(mean - fi) / fi > 0.2 else mean                         # :203-204
    return out


# ---------------------------------------------------------------------------------------------
# Synthesis (W/src/synthesis.cpp)
# ---------------------------------------------------------------------------------------------
def minimum_phase(logspec_half, N):
    """W/src/common.cpp:182-220 for bins 0..N/2 (the imaginary cepstrum is rounding noise)."""
    half = N // 2
    cep = np.fft.rfft(np.concatenate([logspec_half, logspec_half[half - 1:0:-1]])).real
    fold = np.zeros(N)
    fold[0], fold[1:half], fold[half] = cep[0], 2.0 * cep[1:half], cep[half]
    S = np.fft.fft(fold)[:half + 1]
    return np.exp(S.real / N) * (np.cos(S.imag / N) + 1j * np.sin(S.imag / N))


def time_base(f0, fs, frame_period_ms, y_length, lowest_f0):
    """GetTimeBase :287-320 -> (pulse index, time shift, vuv per sample)."""
    n = len(f0)
    fp = frame_period_ms / 1000.0
    cf0 = np.where(f0 < lowest_f0, 0.0, f0)
    cv = np.where(cf0 == 0.0, 0.0, 1.0)
    ct = np.arange(n + 1) * fp
    cf0 = np.append(cf0, cf0[-1] * 2 - cf0[-2])
    cv = np.append(cv, cv[-1] * 2 - cv[-2])
    ta = np.arange(y_length) / float(fs)
    vuv = interp1(ct, cv, ta) > 0.5
    fi = np.where(vuv, interp1(ct, cf0, ta), kDefaultF0)
    total, wrap = np.zeros(y_length), np.zeros(y_length)
    dp = C.POINTER(C.c_double)
    port().port_phase_scan(np.ascontiguousarray(fi).ctypes.data_as(dp), C.c_longlong(y_length), C.c_int(fs),
                           total.ctypes.data_as(dp), wrap.ctypes.data_as(dp))
    idx = np.nonzero(np.abs(np.diff(wrap)) > kPi)[0]
    y1, y2 = wrap[idx] - 2.0 * kPi, wrap[idx + 1]
    return idx, (-y1 / (y2 - y1)) / fs, vuv


def synthesis(f0, sp, ap, fft_size, frame_period_ms, fs, y_length=None):
    """:338-397."""
    N, half = fft_size, fft_size // 2
    if y_length is None:
        y_length = int((len(f0) - 1) * frame_period_ms / 1000.0 * fs) + 1
    y = np.zeros(y_length)
    lowest_f0 = fs // N + 1.0                          # integer division, :359
    idx, shift, vuv = time_base(f0, fs, frame_period_ms, y_length, lowest_f0)
    if len(idx) == 0:
        return y
    rn = randn_stream(int(idx[-1] - idx[0]) + 1)
    rem = 0.5 - 0.5 * np.cos(2.0 * kPi * (np.arange(half) + 1.0) / (1.0 + N))     # GetDCRemover :322-334
    rem = np.concatenate([rem, rem[::-1]])
    rem = rem / rem[:half].sum() / 2.0
    fp = frame_period_ms / 1000.0
    for p, (ix, sh) in enumerate(zip(idx, shift)):
        noise_size = int(idx[min(len(idx) - 1, p + 1)] - ix)                      # :370-371
        pos = (ix / float(fs)) / fp
        a, b = min(len(f0) - 1, int(math.floor(pos))), min(len(f0) - 1, int(math.ceil(pos)))
        w = pos - a
        clamp = lambda v: np.clip(v, 0.001, 0.999999999999)                      # W/src/world/common.h:111-113
        if a == b:
            se, ar = np.abs(sp[a]), clamp(ap[a]) ** 2                            # :140-178
        else:
            se = (1.0 - w) * np.abs(sp[a]) + w * np.abs(sp[b])
            ar = ((1.0 - w) * clamp(ap[a]) + w * clamp(ap[b])) ** 2
        voiced = bool(vuv[ix])
        periodic = np.zeros(N)
        if voiced and not ar[0] > 0.999:                                          # GetPeriodicResponse :105-138
            M = minimum_phase(np.log(se * (1.0 - ar) + kMySafeGuardMinimum) / 2.0, N)
            c = 2.0 * kPi * sh * fs / N
            re2 = np.cos(c * np.arange(half + 1))
            M = M * (re2 - 1j * np.sqrt(1.0 - re2 * re2))                         # :88-100
            M[0], M[half] = M[0].real, M[half].real
            r = np.fft.fftshift(np.fft.irfft(M, N) * N)
            dc = r[half:].sum()                                                   # RemoveDCComponent :73-82
            periodic = np.concatenate([-dc * rem[:half], r[half:] - dc * rem[half:]])
        # GetAperiodicResponse :38-68
        nz = np.zeros(N)
        if noise_size > 0:
            v = rn[ix - idx[0]:ix - idx[0] + noise_size]
            nz[:noise_size] = v - v.sum() / noise_size
        M = minimum_phase(np.log(se * ar) / 2.0 if voiced else np.log(se) / 2.0, N)
        A = np.fft.rfft(nz) * M
        A[0], A[half] = A[0].real, A[half].real
        aper = np.fft.fftshift(np.fft.irfft(A, N) * N)
        resp = (periodic * math.sqrt(noise_size) + aper) / N                      # :214-217
        o0 = ix - half + 1
        lo, hi = max(0, o0), min(y_length, o0 + N)
        y[lo:hi] += resp[lo - o0:hi - o0]
    return y


# ---------------------------------------------------------------------------------------------
# codec (W/src/codec.cpp)
# ---------------------------------------------------------------------------------------------
def code_spectral_envelope(sp, fs, fft_size, ndim):
    """:266-295 (DCTForCodec :72-88, GetParametersForCoding :161-179)."""
    md = fft_size // 2
    mel = lambda f: kM0 * np.log(f / kF0mel + 1.0)
    floor_mel, ceil_mel = mel(40.0), mel(min(fs / 2.0, 20000.0))
    mel_axis = (ceil_mel - floor_mel) * np.arange(md) / md + floor_mel
    knots = mel(np.arange(md) * fs / fft_size)
    i = np.arange(ndim)
    w = 2.0 * (np.cos(i * kPi / fft_size) + 1j * np.sin(i * kPi / fft_size)) / math.sqrt(fft_size)
    w[0] = w[0].real / math.sqrt(2.0) + 1j * w[0].imag
    out = np.zeros((len(sp), ndim))
    for f, row in enumerate(sp):
        m = interp1(knots, np.log(row[:md]), mel_axis)
        wave = np.concatenate([m[0::2], m[::-1][0::2]])
        X = np.fft.rfft(wave)[:ndim]
        out[f] = (X.real * w.real - X.imag * w.imag) / math.sqrt(md)
    return out

"""Synthetic speech-like utterances: the workload every parity test and
bench line runs on.  The reference ships no corpus and defines no generator; this is the
concrete definition used everywhere in this repository.

Per utterance u (all parameters from numpy default_rng(seed=u), so any utterance can be
regenerated independently on any rank):
  * duration T ~ U(1.5, 6.0) s unless given,
  * base F0 log-uniform in [80, 300] Hz, +-15 % sinusoidal modulation at 0.3-1.5 Hz,
  * 40 harmonics (those below 0.45 fs), amplitude 1/k shaped by 3 formant resonances,
  * alternating voiced / unvoiced segments (about 60 % voiced, 10 ms raised-cosine ramps),
  * white Gaussian noise at -34 dB re peak everywhere (no digital silence),
  * peak-normalised to 0.6, quantised to int16; x = int16 / 32768 exactly as the
    reference's wavread does (externs/WORLD_v2/test/audioio.cpp:229-251).
"""
import math
import numpy as np
import torch


def utterance_params(u, duration=None):
    rng = np.random.default_rng(int(u))
    T = float(rng.uniform(1.5, 6.0)) if duration is None else float(duration)
    p = dict(T=T)
    p["f0_base"] = float(math.exp(rng.uniform(math.log(80.0), math.log(300.0))))
    p["mod_rate"] = float(rng.uniform(0.3, 1.5))
    p["mod_phase"] = float(rng.uniform(0, 2 * math.pi))
    p["formants"] = [float(rng.uniform(300, 900)), float(rng.uniform(1000, 2400)),
                     float(rng.uniform(2500, 3800))]
    p["bandwidths"] = [float(rng.uniform(60, 140)), float(rng.uniform(80, 200)),
                       float(rng.uniform(120, 300))]
    # voiced / unvoiced boundaries (seconds); start state random
    segs = []
    t = 0.0
    voiced = bool(rng.uniform() < 0.6)
    while t < T:
        d = float(rng.uniform(0.35, 0.95)) if voiced else float(rng.uniform(0.2, 0.6))
        segs.append((t, min(T, t + d), voiced))
        t += d
        voiced = not voiced
    p["segments"] = segs
    p["noise_seed"] = int(rng.integers(0, 2 ** 31 - 1))
    return p


def _formant_gain(freq, formants, bandwidths):
    g = torch.zeros_like(freq)
    for fc, bw in zip(formants, bandwidths):
        g = g + 1.0 / torch.sqrt(1.0 + ((freq - fc) / bw) ** 2)
    return g + 0.05


def make_utterance(u, fs, duration=None, device="cpu"):
    """-> (pcm int16 tensor [L] on `device`, params dict).  x = pcm / 32768."""
    p = utterance_params(u, duration)
    L = int(round(p["T"] * fs))
    dev = torch.device(device)
    t = torch.arange(L, dtype=torch.float64, device=dev) / fs
    f0 = p["f0_base"] * (1.0 + 0.15 * torch.sin(2 * math.pi * p["mod_rate"] * t + p["mod_phase"]))
    theta = 2 * math.pi * torch.cumsum(f0, 0) / fs
    # voiced gate with 10 ms raised-cosine ramps
    gate = torch.zeros(L, dtype=torch.float64, device=dev)
    ramp = max(1, int(0.010 * fs))
    for (a, b, v) in p["segments"]:
        if not v:
            continue
        ia, ib = int(a * fs), min(L, int(b * fs))
        if ib - ia < 4 * ramp:
            continue
        seg = torch.ones(ib - ia, dtype=torch.float64, device=dev)
        w = 0.5 - 0.5 * torch.cos(math.pi * torch.arange(ramp, dtype=torch.float64, device=dev) / ramp)
        seg[:ramp] = w
        seg[-ramp:] = torch.flip(w, [0])
        gate[ia:ib] = seg
    # harmonic sum by the Chebyshev recurrence sin(k th) = 2 cos(th) sin((k-1) th) - sin((k-2) th)
    c2 = 2.0 * torch.cos(theta)
    s_prev = torch.zeros_like(theta)
    s_cur = torch.sin(theta)
    voiced = torch.zeros_like(theta)
    for k in range(1, 41):
        fk = k * f0
        amp = _formant_gain(fk, p["formants"], p["bandwidths"]) / k
        amp = torch.where(fk < 0.45 * fs, amp, torch.zeros_like(amp))
        voiced = voiced + amp * s_cur
        s_prev, s_cur = s_cur, c2 * s_cur - s_prev
    sig = gate * voiced
    peak = float(sig.abs().max().item()) if L else 1.0
    peak = peak if peak > 0 else 1.0
    noise = np.random.default_rng(p["noise_seed"]).standard_normal(L)
    noise = torch.from_numpy(noise).to(dev)
    sig = sig / peak + (10.0 ** (-34.0 / 20.0)) * noise
    sig = sig * (0.6 / float(sig.abs().max().item()))
    pcm = torch.clamp(torch.round(sig * 32767.0), -32768, 32767).to(torch.int16)
    return pcm, p


def make_corpus(n_utt, fs, first=0, duration=None, device="cpu"):
    """List of int16 tensors (one per utterance).  Utterance ids first .. first+n_utt-1."""
    return [make_utterance(first + i, fs, duration, device)[0] for i in range(n_utt)]


def pcm_to_double(pcm):
    """wavread's convention: int16 / 32768.0 (W/test/audioio.cpp:229-251)."""
    if isinstance(pcm, torch.Tensor):
        pcm = pcm.cpu().numpy()
    return pcm.astype(np.float64) / 32768.0

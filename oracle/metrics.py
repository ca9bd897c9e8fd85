"""TEST INFRASTRUCTURE ONLY — the five parity metrics of BASELINE.json's north_star,
as defined in SURVEY.md §8(c).  All inputs are float64 numpy arrays; `ref` is the
oracle's output and `new` is ours on identical inputs."""
import numpy as np

# tolerances stated by north_star
TOL_VUV_AGREEMENT = 0.999     # fraction of frames
TOL_F0_REL = 1e-4             # voiced F0 relative error (max over frames voiced in both)
TOL_LSD_DB = 0.01             # log spectral distance, max over frames
TOL_AP_ABS = 1e-4             # aperiodicity absolute error, max over frames and bins
TOL_SNR_DB = 60.0             # resynthesis SNR


def vuv_agreement(f0_ref, f0_new):
    if len(f0_ref) == 0:
        return 1.0
    return float(np.mean((f0_ref > 0) == (f0_new > 0)))


def f0_rel_error(f0_ref, f0_new):
    both = (f0_ref > 0) & (f0_new > 0)
    if not both.any():
        return 0.0
    return float(np.max(np.abs(f0_new[both] - f0_ref[both]) / f0_ref[both]))


def lsd_db(sp_ref, sp_new):
    """per-frame log spectral distance in dB -> (mean, max)."""
    if sp_ref.size == 0:
        return 0.0, 0.0
    d = 10.0 * np.log10(sp_new / sp_ref)
    per_frame = np.sqrt(np.mean(d * d, axis=-1))
    return float(per_frame.mean()), float(per_frame.max())


def ap_abs_error(ap_ref, ap_new):
    if ap_ref.size == 0:
        return 0.0
    return float(np.max(np.abs(ap_new - ap_ref)))


def snr_db(y_ref, y_new):
    err = float(np.sum((y_ref - y_new) ** 2))
    sig = float(np.sum(y_ref ** 2))
    if err == 0.0:
        return float("inf")
    return 10.0 * np.log10(sig / err)

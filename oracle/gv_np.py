"""TEST INFRASTRUCTURE (oracle): global-variance statistics of the static feature streams.

Restates, in numpy, what the reference does with SPTK's `vstat -d -o 2` (SPTK is an external
dependency of the reference, not vendored in /root/reference; its published algorithm: per-dimension
sum and sum of squares of the float32 input accumulated in double in input order, mean = sum / k,
variance = sumsq / k - mean^2 over the k vectors read, float32 output):

  * scripts/Training.pl make_data_gv :1402-1456 -- for every utterance and every stream of @cmp
    (mgc, lf0, bap): the variance of each dimension over the frames of the utterance; MSD streams
    (lf0) first lose their unvoiced frames (`grep -v '1e+10'`, :1437); the vectors are concatenated
    into the utterance's GV observation;
  * data/Makefile.in:447-458 -- `vstat` once more over the per-utterance variance vectors (a
    float32 file, tmp.var1): the variance of the variances, stats/gv.var.

Only tests/ may import this module.  Parity is pinned to the formulas above (the reference holds no
golden vector for them); tests/test_gv.py checks the restatement against a direct two-pass
evaluation and the CUDA path against the restatement."""
import numpy as np


def vstat_diag_var(x):
    """`vstat -d -o 2` of a [k][dim] float32 array -> float64 variances (before the float32 output)."""
    x = np.asarray(x, np.float32)
    if x.ndim == 1:
        x = x[:, None]
    k = x.shape[0]
    if k == 0:
        return np.full(x.shape[1], np.nan)
    s = np.zeros(x.shape[1])
    q = np.zeros(x.shape[1])
    for row in x.astype(np.float64):          # input order, double accumulators
        s += row
        q += row * row
    mean = s / k
    return q / k - mean * mean


def utterance_gv(mgc, lf0, bap):
    """[mgc | lf0 (voiced frames) | bap] variances of one utterance (lf0 == 0 marks unvoiced here)."""
    lf0 = np.asarray(lf0, np.float32)
    return np.concatenate([vstat_diag_var(mgc), vstat_diag_var(lf0[lf0 != 0]), vstat_diag_var(bap)])


def corpus_gv(per_utt):
    """mean and variance over utterances of the float32-rounded per-utterance variances (NaN rows of a
    column are skipped, as the reference drops utterances whose GV observation holds a NaN)."""
    per_utt = np.asarray(per_utt, np.float64)
    mean = np.zeros(per_utt.shape[1])
    var = np.zeros(per_utt.shape[1])
    for c in range(per_utt.shape[1]):
        v = per_utt[:, c]
        v = v[~np.isnan(v)].astype(np.float32)
        if len(v) == 0:
            mean[c] = var[c] = np.nan
            continue
        vv = vstat_diag_var(v)
        mean[c] = float(np.sum(v.astype(np.float64)) / len(v))
        var[c] = vv[0]
    return mean, var

"""Corpus-level host logic: utterance sharding across ranks and the lf0 statistics reduce.

The reference runs one `analysis` process per utterance in a serial shell loop
(data/Makefile.in:125-242) and later computes corpus statistics with SPTK `vstat`
(data/Makefile.in:414-459).  Here utterances are dealt to the GPUs of one box; the hot path
has no collective; one all-reduce of {count, sum, sum of squares} of voiced log-f0 per rank
forms the corpus mean / variance."""
import numpy as np


def shard_utterances(lengths, rank, world_size):
    """Longest-first greedy partition by sample count -> sorted utterance ids of `rank`.
    Deterministic, identical on every rank, no communication."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, np.int64)
    owner = np.zeros(len(lengths), np.int32)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += int(lengths[i])
    return np.nonzero(owner == rank)[0]


def merge_stats(partials):
    """partials: iterable of [count, sum, sumsq] -> dict(count, mean, var)."""
    tot = np.sum(np.asarray(list(partials), np.float64), axis=0)
    n = tot[0]
    if n <= 0:
        return dict(count=0.0, mean=0.0, var=0.0)
    mean = tot[1] / n
    return dict(count=float(n), mean=float(mean), var=float(max(0.0, tot[2] / n - mean * mean)))


def allreduce_stats(local3, group=None):
    """Sum [count, sum, sumsq] over all ranks with torch.distributed (NCCL on GPUs, gloo on
    CPU).  Returns the merged dict.  Without an initialised process group it is the identity."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(local3, np.float64))
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t.cpu()
    return merge_stats([t.numpy()])


def merge_gv(partials):
    """partials: iterable of [dims][3] = {count, sum, sum of squares} of the per-utterance variances
    (wb200_batch_gv_stats, one per batch / rank) -> (mean[dims], var[dims]): the mean and the variance
    of the variances over the corpus (stats/gv.var of data/Makefile.in:447-458)."""
    tot = np.sum(np.asarray(list(partials), np.float64), axis=0)
    n = np.maximum(tot[:, 0], 1.0)
    mean = tot[:, 1] / n
    var = tot[:, 2] / n - mean * mean
    bad = tot[:, 0] <= 0
    mean[bad] = np.nan
    var[bad] = np.nan
    return mean, var

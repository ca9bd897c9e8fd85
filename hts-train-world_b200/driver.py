"""Corpus driver: the `features:` target of the reference's data/Makefile.in (lines 121-242)
for the WORLD branch, without the per-utterance process fork.

The reference loops over raw/*.raw in a shell `for`, and for every file: clip check with
x2x|minmax (:127-129), raw2wav (:213), `$(WORLD)/analysis wav lf0 mgc bap FRAMEPERIOD FFTLEN
MGCDIM` (:214), then SPTK `nan` checks that delete an output containing NaN (:216-238).  Here
the utterances of a shard are read, batched by audio duration, analysed on the GPU in one go
per batch and written as the same float32 lf0 / mgc / bap files; the statistics partials
({count, sum, sum of squares} of voiced lf0 and of every mgc dimension) are accumulated and
all-reduced over the ranks at the end.

  python hts-train-world_b200/driver.py --raw-dir data/raw --out-dir data --fs 48000 [--f0 harvest]
  torchrun --nproc-per-node 8 hts-train-world_b200/driver.py ...     # utterance-sharded, NCCL stats

With --cmp the `cmp:` target (:244-321) runs in the same pass while the features are still in
HBM: every stream is extended by its delta windows (data/scripts/window.pl, data/win/*.win[123]),
the streams are merged side by side (mgc | lf0 | bap) and written with the 12-byte HTK header of
data/scripts/addhtkheader.pl as cmp/<base>.cmp; the per-column {count, sum, sum of squares} of the
cmp matrix join the all-reduce.  Extract.py's label-driven streams (the second lf0 dimension and
vib, :215) are singing-voice specific and stay outside; a caller that has them passes
them to Corpus.compose_cmp as host streams.
"""
import argparse
import glob
import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))


def fftlen_for(fs):
    """configure.ac:540-549 (equals GetFFTSizeForCheapTrick at f0_floor 71 Hz)."""
    n = 512
    while n * 25 < fs:          # 1024 for fs <= 25.6 kHz, 2048 for <= 51.2 kHz, ...
        n *= 2
    return max(1024, n)


def read_raw(path):
    return np.fromfile(path, dtype="<i2")


def passes_clip_check(pcm):
    """data/Makefile.in:127-129: non-empty, min > -32768 and max < 32767."""
    return pcm.size > 0 and int(pcm.min()) > -32768 and int(pcm.max()) < 32767


def extract_features(raw_paths, out_dir, fs=48000, frame_period_ms=5.0, mgc_dim=50, bap_dim=24,
                     f0="dio", batch_seconds=4000.0, rank=0, world=1, log=print, cmp=False, windows=None):
    """Returns dict(done=[...], skipped=[...], failed={base: [streams]}, stats=[(1+mgc_dim), 3]
    and, with cmp=True, cmp_stats=[cmp_dim, 3])."""
    import hts_train_world_b200 as wb
    from hts_train_world_b200 import corpus
    for d in ("lf0", "mgc", "bap") + (("cmp",) if cmp else ()):
        os.makedirs(os.path.join(out_dir, d), exist_ok=True)
    frame_shift = int(round(frame_period_ms * fs / 1000.0))     # FRAMESHIFT in samples
    cmp_stats = None
    sizes = [os.path.getsize(p) // 2 for p in raw_paths]
    mine = corpus.shard_utterances(sizes, rank, world)
    report = dict(done=[], skipped=[], failed={})
    stats = np.zeros((1 + mgc_dim, 3))
    batch, batch_audio = [], 0.0

    def flush():
        nonlocal batch, batch_audio, cmp_stats
        if not batch:
            return
        names, pcms = zip(*batch)
        c = wb.Corpus(fs, [len(p) for p in pcms], frame_period_ms)
        c.upload_pcm16(np.concatenate(pcms))
        c.analyze(f0=f0)
        c.code(mgc_dim, bap_dim)
        lf0, mgc, bap = c.coded()
        stats[:] += c.feature_stats()
        cmp_rows = None
        if cmp:
            cmp_rows = c.compose_cmp(("mgc", "lf0", "bap"), windows)
            st = c.cmp_stats()
            cmp_stats = st if cmp_stats is None else cmp_stats + st
        for u, base in enumerate(names):
            sl = c.frames_of(u)
            bad = []
            for stream, arr in (("lf0", lf0[sl]), ("mgc", mgc[sl]), ("bap", bap[sl])):
                if np.isnan(arr).any():                 # the SPTK `nan` checks, :216-238
                    log(" Failed to extract features from %s: %s error" % (base, stream.upper()))
                    bad.append(stream)
                    continue
                arr.astype("<f4").tofile(os.path.join(out_dir, stream, "%s.%s" % (base, stream)))
            if bad:
                report["failed"][base] = bad
            elif cmp:                                    # data/Makefile.in:284: only when every stream exists
                rows = cmp_rows[sl]
                with open(os.path.join(out_dir, "cmp", base + ".cmp"), "wb") as f:
                    f.write(wb.htk_header(rows.shape[0], fs, frame_shift, 4 * rows.shape[1], 9))
                    f.write(rows.astype("<f4").tobytes())
            report["done"].append(base)
        c.close()
        batch, batch_audio = [], 0.0

    for i in mine:
        path = raw_paths[i]
        base = os.path.splitext(os.path.basename(path))[0]
        pcm = read_raw(path)
        if not passes_clip_check(pcm):
            report["skipped"].append(base)
            continue
        log("Extracting features from %s" % path)
        batch.append((base, pcm))
        batch_audio += len(pcm) / float(fs)
        if batch_audio >= batch_seconds:
            flush()
    flush()
    import torch.distributed as dist
    if cmp and cmp_stats is None:                        # a rank without utterances still joins the reduce
        dims = [(mgc_dim, 0), (1, 1), (bap_dim, 2)]
        cmp_stats = np.zeros((sum(d * (3 if windows is None else len(windows[i])) for d, i in dims), 3))
    if dist.is_available() and dist.is_initialized():
        import torch
        parts = [stats] + ([cmp_stats] if cmp else [])
        t = torch.as_tensor(np.concatenate(parts))       # one all-reduce for all partials
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        t = t.cpu().numpy()
        stats = t[:stats.shape[0]]
        if cmp:
            cmp_stats = t[stats.shape[0]:]
    report["stats"] = stats
    if cmp:
        report["cmp_stats"] = cmp_stats
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw-dir", required=True)
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--fs", type=int, default=48000)
    ap.add_argument("--frameshift", type=int, default=None, help="samples (FRAMESHIFT); default 5 ms")
    ap.add_argument("--mgc-order", type=int, default=49)
    ap.add_argument("--bap-dim", type=int, default=24)
    ap.add_argument("--f0", default="dio", choices=["dio", "harvest"])
    ap.add_argument("--cmp", action="store_true", help="also compose cmp/<base>.cmp (delta windows + HTK header)")
    ap.add_argument("--win-dir", default=None, help="directory with {mgc,lf0,bap}.win[123] (default: static, delta, delta-delta)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import hts_train_world_b200 as wb
    from hts_train_world_b200 import corpus
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wb.init(local)
    fp = 5.0 if args.frameshift is None else args.frameshift * 1000.0 / args.fs
    paths = sorted(glob.glob(os.path.join(args.raw_dir, "*.raw")))
    windows = None
    if args.win_dir:
        windows = [tuple(wb.parse_window_file(open(os.path.join(args.win_dir, "%s.win%d" % (s, i))).read())
                         for i in (1, 2, 3) if os.path.exists(os.path.join(args.win_dir, "%s.win%d" % (s, i))))
                   for s in ("mgc", "lf0", "bap")]
    rep = extract_features(paths, args.out_dir, args.fs, fp, args.mgc_order + 1, args.bap_dim, args.f0,
                           rank=rank, world=world, log=(print if rank == 0 else (lambda *_: None)),
                           cmp=args.cmp, windows=windows)
    if rank == 0:
        st = rep["stats"]
        summary = dict(lf0=corpus.merge_stats([st[0]]), mgc=[corpus.merge_stats([r]) for r in st[1:]])
        if args.cmp:
            summary["cmp"] = [corpus.merge_stats([r]) for r in rep["cmp_stats"]]
        json.dump(summary, open(os.path.join(args.out_dir, "world_b200_stats.json"), "w"), indent=1)
        print("done: %d utterances on rank 0, %d skipped (clip check), %d with NaN streams"
              % (len(rep["done"]), len(rep["skipped"]), len(rep["failed"])))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

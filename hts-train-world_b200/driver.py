"""Corpus driver: the `features:` target of the reference's data/Makefile.in (lines 121-242)
for the WORLD branch, without the per-utterance process fork.

The reference loops over raw/*.raw in a shell `for`, and for every file: clip check with
x2x|minmax (:127-129), raw2wav (:213), `$(WORLD)/analysis wav lf0 mgc bap FRAMEPERIOD FFTLEN
MGCDIM` (:214), then SPTK `nan` checks that delete an output containing NaN (:216-238).  Here
the utterances of a shard are read, batched by audio duration, analysed on the GPU in one go
per batch and written as the same float32 lf0 / mgc / bap files; the statistics partials
({count, sum, sum of squares} of voiced lf0 and of every mgc dimension, and the global-variance
partials of scripts/Training.pl:1402-1456 / data/Makefile.in:447-458) are accumulated and
all-reduced over the ranks at the end.

  python hts-train-world_b200/driver.py --raw-dir data/raw --out-dir data --fs 48000 [--f0 harvest]
  torchrun --nproc-per-node 8 hts-train-world_b200/driver.py ...     # utterance-sharded, NCCL stats
  python hts-train-world_b200/driver.py --synthetic-hours 100 --out-dir /tmp/out ...   # BASELINE config 5

The job is a three-stage pipeline: a reader thread loads and batches the next utterances while the
GPU works on the current batch (its PCM goes up with the asynchronous upload, the coded features
come back with asynchronous copies into one of two sets of pinned host buffers), and a writer
thread turns the previous batch's buffers into files.  --resume skips utterances whose outputs
already exist (their files still enter the statistics, read back on the host).

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
import queue
import sys
import threading
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

STREAMS = ("lf0", "mgc", "bap")


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


def outputs_exist(out_dir, base, cmp):
    """--resume: every output file of the utterance is there and not empty."""
    names = [os.path.join(out_dir, s, "%s.%s" % (base, s)) for s in STREAMS]
    if cmp:
        names.append(os.path.join(out_dir, "cmp", base + ".cmp"))
    return all(os.path.exists(n) and os.path.getsize(n) > 0 for n in names)


def raw_source(paths):
    """(base, int16 array) of every raw file, in order."""
    for path in paths:
        yield os.path.splitext(os.path.basename(path))[0], read_raw(path)


SYNTH_POOL = 256


def synthetic_pool(ids, fs, pool=SYNTH_POOL, device="cuda"):
    """BASELINE config 5: the signals of the synthetic corpus (signals.py).  Generating 100 hours of distinct
    speech-like signals would cost more than analysing them, so the corpus is drawn from a pool of `pool`
    distinct utterances generated once on the device (utterance u is signal u mod pool).  Producing the
    input is not part of the job that is timed: main() builds the pool before it starts the clock."""
    from hts_train_world_b200 import signals
    need = sorted(set(int(u) % pool for u in ids))
    return {k: signals.make_utterance(k, fs, device=device)[0].cpu().numpy() for k in need}


def synthetic_source(ids, cache, pool=SYNTH_POOL):
    """(base, int16 array) of every utterance of this rank: each one is still uploaded, analysed, coded and
    written on its own."""
    for u in ids:
        yield "synth_%07d" % int(u), cache[int(u) % pool]


class _Slot:
    """One set of pinned host buffers for the results of a batch (two of them alternate)."""

    def __init__(self):
        self.cap = 0
        self.free = threading.Event()
        self.free.set()

    def ensure(self, frames, mgc_dim, bap_dim, cmp_dim):
        import torch
        if frames > self.cap or getattr(self, "cmp_dim", -1) != cmp_dim:
            self.cap = int(frames * 1.25) + 1024
            self.cmp_dim = cmp_dim
            self.lf0 = torch.empty(self.cap, dtype=torch.float32).pin_memory()
            self.mgc = torch.empty((self.cap, mgc_dim), dtype=torch.float32).pin_memory()
            self.bap = torch.empty((self.cap, bap_dim), dtype=torch.float32).pin_memory()
            self.cmp = torch.empty((self.cap, cmp_dim), dtype=torch.float32).pin_memory() if cmp_dim else None


def extract_features(source, out_dir, fs=48000, frame_period_ms=5.0, mgc_dim=50, bap_dim=24,
                     f0="dio", batch_seconds=4000.0, log=print, cmp=False, windows=None, resume=False,
                     write_files=True, reduce=True):
    """source: iterable of (base, int16 pcm) for the utterances of THIS rank.
    Returns dict(done, skipped, resumed, failed, stats[(1+mgc_dim), 3], gv[(mgc_dim+1+bap_dim), 3],
    cmp_stats (with cmp=True), seconds=dict(...))."""
    import torch
    import hts_train_world_b200 as wb
    for d in STREAMS + (("cmp",) if cmp else ()):
        os.makedirs(os.path.join(out_dir, d), exist_ok=True)
    frame_shift = int(round(frame_period_ms * fs / 1000.0))     # FRAMESHIFT in samples
    report = dict(done=[], skipped=[], resumed=[], failed={})
    stats = np.zeros((1 + mgc_dim, 3))
    gv = np.zeros((mgc_dim + 1 + bap_dim, 3))
    cmp_stats = [None]
    timers = dict(read=0.0, gpu=0.0, write=0.0, audio=0.0)
    t_start = time.perf_counter()

    # ---- stage 1: reader thread (file I/O, clip check, resume check, batching) --------------------------
    batches = queue.Queue(maxsize=2)

    def reader():
        batch, audio = [], 0.0
        t0 = time.perf_counter()
        for base, pcm in source:
            if not passes_clip_check(pcm):
                report["skipped"].append(base)
                continue
            if resume and outputs_exist(out_dir, base, cmp):
                report["resumed"].append(base)
                continue
            batch.append((base, pcm))
            audio += len(pcm) / float(fs)
            if audio >= batch_seconds:
                timers["read"] += time.perf_counter() - t0
                batches.put(_pack(batch))
                t0 = time.perf_counter()
                batch, audio = [], 0.0
        if batch:
            batches.put(_pack(batch))
        timers["read"] += time.perf_counter() - t0
        batches.put(None)

    def _pack(batch):
        names = [b for b, _ in batch]
        lengths = [len(p) for _, p in batch]
        pcm = torch.empty(sum(lengths), dtype=torch.int16).pin_memory()
        o = 0
        view = pcm.numpy()
        for _, p in batch:
            view[o:o + len(p)] = p
            o += len(p)
        return names, lengths, pcm

    # ---- stage 3: writer thread (NaN checks, files) ------------------------------------------------------
    to_write = queue.Queue()

    def writer():
        while True:
            job = to_write.get()
            if job is None:
                return
            t0 = time.perf_counter()
            names, slices, slot = job
            lf0, mgc, bap = slot.lf0.numpy(), slot.mgc.numpy(), slot.bap.numpy()
            rows_all = slot.cmp.numpy() if cmp else None
            for base, sl in zip(names, slices):
                bad = []
                for stream, arr in (("lf0", lf0[sl]), ("mgc", mgc[sl]), ("bap", bap[sl])):
                    path = os.path.join(out_dir, stream, "%s.%s" % (base, stream))
                    if np.isnan(arr).any():                 # the SPTK `nan` checks, :216-238: the file is removed
                        log(" Failed to extract features from %s: %s error" % (base, stream.upper()))
                        bad.append(stream)
                        if os.path.exists(path):
                            os.remove(path)
                        continue
                    if write_files:
                        arr.astype("<f4", copy=False).tofile(path)
                if bad:
                    report["failed"][base] = bad
                    stale = os.path.join(out_dir, "cmp", base + ".cmp")
                    if cmp and os.path.exists(stale):
                        os.remove(stale)
                    continue
                if cmp and write_files:                      # data/Makefile.in:284: only when every stream exists
                    rows = rows_all[sl]
                    with open(os.path.join(out_dir, "cmp", base + ".cmp"), "wb") as f:
                        f.write(wb.htk_header(rows.shape[0], fs, frame_shift, 4 * rows.shape[1], 9))
                        f.write(rows.astype("<f4", copy=False).tobytes())
                report["done"].append(base)
            slot.free.set()
            timers["write"] += time.perf_counter() - t0

    th_r = threading.Thread(target=reader, daemon=True)
    th_w = threading.Thread(target=writer, daemon=True)
    th_r.start()
    th_w.start()

    # ---- stage 2: this thread drives the GPU ----------------------------------------------------------------
    slots = [_Slot(), _Slot()]
    prev = None                                             # (corpus, names, slices, slot) whose copies are in flight
    n_batch = 0
    while True:
        item = batches.get()
        if item is None:
            break
        names, lengths, pcm = item
        t0 = time.perf_counter()
        c = wb.Corpus(fs, lengths, frame_period_ms)
        c.upload_pcm16_async(pcm)
        c.analyze(f0=f0)
        c.code(mgc_dim, bap_dim)
        stats += c.feature_stats()
        gv += c.gv_stats()[1]
        if cmp:
            c.compose_cmp(("mgc", "lf0", "bap"), windows, fetch=False)
            st = c.cmp_stats()
            cmp_stats[0] = st if cmp_stats[0] is None else cmp_stats[0] + st
        slot = slots[n_batch % 2]
        slot.free.wait()                                    # the writer is done with this buffer set
        slot.free.clear()
        slot.ensure(c.total_frames, mgc_dim, bap_dim, c.cmp_dim if cmp else 0)
        F = c.total_frames
        c.coded_async(slot.lf0[:F], slot.mgc[:F], slot.bap[:F])
        if cmp:
            c.cmp_async(slot.cmp[:F])
        timers["audio"] += sum(lengths) / float(fs)
        if prev is not None:                                # the previous batch's copies ran beside this batch's kernels
            prev[0].wait_downloads()
            to_write.put(prev[1:])
            prev[0].close()
        prev = (c, names, [c.frames_of(u) for u in range(len(names))], slot)
        n_batch += 1
        timers["gpu"] += time.perf_counter() - t0
    if prev is not None:
        prev[0].wait_downloads()
        to_write.put(prev[1:])
        prev[0].close()
    to_write.put(None)
    th_w.join()
    th_r.join()

    # (utterances skipped by --resume still belong to the corpus statistics: main() adds stats_from_files)
    report["stats"], report["gv"] = stats, gv
    if cmp:
        report["cmp_stats"] = cmp_stats[0]
    timers["total"] = time.perf_counter() - t_start
    report["seconds"] = timers
    return report


def stats_from_files(out_dir, bases, mgc_dim, bap_dim):
    """{count, sum, sum of squares} partials (feature statistics and global variance) of utterances whose
    files already exist (--resume): the same quantities the device kernels deliver, from the float32 files."""
    stats = np.zeros((1 + mgc_dim, 3))
    gv = np.zeros((mgc_dim + 1 + bap_dim, 3))
    for base in bases:
        lf0 = np.fromfile(os.path.join(out_dir, "lf0", base + ".lf0"), "<f4").astype(np.float64)
        mgc = np.fromfile(os.path.join(out_dir, "mgc", base + ".mgc"), "<f4").astype(np.float64).reshape(-1, mgc_dim)
        bap = np.fromfile(os.path.join(out_dir, "bap", base + ".bap"), "<f4").astype(np.float64).reshape(-1, bap_dim)
        v = lf0[lf0 != 0]
        stats[0] += [len(v), v.sum(), (v * v).sum()]
        stats[1:, 0] += len(mgc)
        stats[1:, 1] += mgc.sum(axis=0)
        stats[1:, 2] += (mgc * mgc).sum(axis=0)
        cols = [mgc[:, i] for i in range(mgc_dim)] + [v] + [bap[:, i] for i in range(bap_dim)]
        for k, col in enumerate(cols):
            if len(col) == 0:
                continue
            var = np.float64(np.float32((col * col).sum() / len(col) - (col.sum() / len(col)) ** 2))
            gv[k] += [1.0, var, var * var]
    return stats, gv


def all_reduce_partials(parts):
    """One all-reduce (NCCL on GPUs, gloo on CPU) of the concatenated partials; identity without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return parts
    import torch
    flat = torch.as_tensor(np.concatenate([np.asarray(p, np.float64).ravel() for p in parts]))
    if dist.get_backend() == "nccl":
        flat = flat.cuda()
    dist.all_reduce(flat)
    flat = flat.cpu().numpy()
    out, o = [], 0
    for p in parts:
        n = int(np.asarray(p).size)
        out.append(flat[o:o + n].reshape(np.asarray(p).shape))
        o += n
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw-dir", default=None)
    ap.add_argument("--synthetic-hours", type=float, default=0.0, help="BASELINE config 5: a synthetic 48 kHz corpus of this many hours instead of raw files")
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--fs", type=int, default=48000)
    ap.add_argument("--frameshift", type=int, default=None, help="samples (FRAMESHIFT); default 5 ms")
    ap.add_argument("--mgc-order", type=int, default=49)
    ap.add_argument("--bap-dim", type=int, default=24)
    ap.add_argument("--f0", default="dio", choices=["dio", "harvest"])
    ap.add_argument("--cmp", action="store_true", help="also compose cmp/<base>.cmp (delta windows + HTK header)")
    ap.add_argument("--win-dir", default=None, help="directory with {mgc,lf0,bap}.win[123] (default: static, delta, delta-delta)")
    ap.add_argument("--resume", action="store_true", help="skip utterances whose output files already exist")
    ap.add_argument("--batch-seconds", type=float, default=4000.0)
    ap.add_argument("--no-files", action="store_true", help="measure without writing the outputs")
    args = ap.parse_args()
    if (args.raw_dir is None) == (args.synthetic_hours <= 0):
        ap.error("give either --raw-dir or --synthetic-hours")
    import torch
    import hts_train_world_b200 as wb
    from hts_train_world_b200 import corpus, signals
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wb.init(local)
    wb.set_copy_deferral(True)      # bulk copies start under D4C's main kernel, not beside the next batch's Dio (world_b200.h)
    fp = 5.0 if args.frameshift is None else args.frameshift * 1000.0 / args.fs
    mgc_dim = args.mgc_order + 1
    if args.raw_dir:
        paths = sorted(glob.glob(os.path.join(args.raw_dir, "*.raw")))
        sizes = [os.path.getsize(p) // 2 for p in paths]
        mine = corpus.shard_utterances(sizes, rank, world)
        source = raw_source([paths[i] for i in mine])
        total_audio = sum(sizes) / float(args.fs)
    else:
        durs, tot = [], 0.0
        while tot < args.synthetic_hours * 3600.0:          # utterance ids 0 .. n-1 of the synthetic corpus
            d = signals.utterance_params(len(durs) % 256)["T"]
            durs.append(d)
            tot += d
        sizes = [int(round(d * args.fs)) for d in durs]
        mine = corpus.shard_utterances(sizes, rank, world)
        source = synthetic_source(mine, synthetic_pool(mine, args.fs))
        total_audio = tot
    windows = None
    if args.win_dir:
        windows = [tuple(wb.parse_window_file(open(os.path.join(args.win_dir, "%s.win%d" % (s, i))).read())
                         for i in (1, 2, 3) if os.path.exists(os.path.join(args.win_dir, "%s.win%d" % (s, i))))
                   for s in ("mgc", "lf0", "bap")]
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                                       # every rank starts the job together
    torch.cuda.synchronize()
    t0 = time.perf_counter()                                 # the job: read, batch, analyse, code, write, reduce
    rep = extract_features(source, args.out_dir, args.fs, fp, mgc_dim, args.bap_dim, args.f0,
                           batch_seconds=args.batch_seconds, log=(print if rank == 0 else (lambda *_: None)),
                           cmp=args.cmp, windows=windows, resume=args.resume, write_files=not args.no_files)
    stats, gv = rep["stats"], rep["gv"]
    if args.resume and rep["resumed"]:
        s2, g2 = stats_from_files(args.out_dir, rep["resumed"], mgc_dim, args.bap_dim)
        stats, gv = stats + s2, gv + g2
    parts = [stats, gv]
    if args.cmp:
        cs = rep["cmp_stats"]
        if cs is None:                                       # a rank without utterances still joins the reduce
            dims = [(mgc_dim, 0), (1, 1), (args.bap_dim, 2)]
            cs = np.zeros((sum(d * (3 if windows is None else len(windows[i])) for d, i in dims), 3))
        parts.append(cs)
    parts = all_reduce_partials(parts)
    wall = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([wall, rep["seconds"]["audio"]], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        wall, audio_done = float(tmax[0]), float(t[1])
    else:
        audio_done = rep["seconds"]["audio"]
    if rank == 0:
        st = parts[0]
        gv_mean, gv_var = corpus.merge_gv([parts[1]])
        summary = dict(lf0=corpus.merge_stats([st[0]]), mgc=[corpus.merge_stats([r]) for r in st[1:]],
                       gv=dict(mean=[float(v) for v in gv_mean], var=[float(v) for v in gv_var],
                               order="mgc[%d] | lf0 | bap[%d]" % (mgc_dim, args.bap_dim)))
        if args.cmp:
            summary["cmp"] = [corpus.merge_stats([r]) for r in parts[2]]
        summary["job"] = dict(ranks=world, audio_seconds=audio_done, corpus_seconds=total_audio, wall_seconds=wall,
                              xRT=audio_done / wall if wall > 0 else 0.0, rank0_seconds=rep["seconds"],
                              files=not args.no_files, f0=args.f0)
        json.dump(summary, open(os.path.join(args.out_dir, "world_b200_stats.json"), "w"), indent=1)
        print("done: %d utterances on rank 0 (%d resumed), %d skipped (clip check), %d with NaN streams; "
              "%.1f s of audio on %d rank(s) in %.2f s wall = %.0f xRT including file I/O"
              % (len(rep["done"]), len(rep["resumed"]), len(rep["skipped"]), len(rep["failed"]),
                 audio_done, world, wall, audio_done / max(wall, 1e-9)))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

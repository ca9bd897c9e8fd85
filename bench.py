#!/usr/bin/env python
"""bench.py — xRT (audio seconds per second) of WORLD analysis + synthesis at 48 kHz / 5 ms.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--utts U]

Workload (BASELINE.json configs[1]): a 1 132-utterance ARCTIC-sized synthetic
48 kHz corpus (hts-train-world_b200/signals.py, seeds = utterance ids), fft_size 2048.  One
"step" = Dio -> StoneMask -> CheapTrick -> D4C -> Synthesis over the whole corpus shard of this
rank, plus the lf0 statistics partials; with N > 1 every rank owns its own 1 132 utterances
(weak scaling, utterance-sharded, no collective on the hot path) and one all-reduce of the
three lf0 statistics closes the step.

  value  whole-job xRT with the int16 PCM already resident in HBM (device-timed, max over ranks)
  e2e    the same through the public batch API from pinned HOST memory: H2D of the PCM, all five
         stages, D2H of the resynthesised 16-bit waveform and of the F0 contour, every step
  roofline      the dominant kernel, timed live with CUDA events on the library stream
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores, bounded sample

--impl reference times the unmodified reference WORLD_v2 (oracle/_ref, built from
/root/reference by oracle/Makefile) on all host cores, one process per core.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 48000
FRAME_PERIOD = 5.0
METRIC = "xRT (audio s/s) WORLD analysis+synthesis, 48 kHz 5 ms"
UNIT = "audio_s/s"
E2E_PARTS = int(os.environ.get("WB_E2E_PARTS", "2"))    # pipelined sub-batches of the end-to-end leg
MGC_DIM, BAP_DIM = 50, 24       # MGCORDER + 1 and the tool's default (data/Makefile.in:214, analysis.cpp:324-328)


# =================================================================================================
# reference arm / cpu baseline: the compiled reference, one process per host core
# =================================================================================================
def _worker_main(conn, opt):
    """Worker process: owns one copy of the reference library (its RNG is a global, so threads
    are not an option: SURVEY.md 5) and a private cache of generated utterances."""
    import torch
    torch.set_num_threads(1)
    from hts_train_world_b200 import signals
    from oracle import ref
    R = ref.load(opt=opt)
    cache = {}
    while True:
        cmd, utts = conn.recv()
        if cmd == "quit":
            return
        if cmd == "prep":
            for u in utts:
                if u not in cache:
                    cache[u] = signals.pcm_to_double(signals.make_utterance(u, FS)[0])
            conn.send(len(cache))
            continue
        # "run": Dio -> StoneMask -> CheapTrick -> D4C -> Synthesis with the analysis tool's options
        # (W/test/analysis.cpp:93-203, W/test/synth.cpp:259) and the tool's codec tail (:293-358); per-stage seconds
        acc = np.zeros(6)
        for u in utts:
            x = cache[u]
            t = [time.perf_counter()]
            tp, f0r = R.dio(x, FS, frame_period=FRAME_PERIOD); t.append(time.perf_counter())
            f0 = R.stonemask(x, FS, tp, f0r); t.append(time.perf_counter())
            sp = R.cheaptrick(x, FS, tp, f0); t.append(time.perf_counter())
            n = sp.shape[1] * 2 - 2
            ap = R.d4c(x, FS, tp, f0, n, threshold=0.0); t.append(time.perf_counter())
            R.synthesis(f0, sp, ap, n, FRAME_PERIOD, FS); t.append(time.perf_counter())
            R.tool_features(f0, sp, ap, FS, n, MGC_DIM, BAP_DIM); t.append(time.perf_counter())
            acc += np.diff(t)
        conn.send(acc)


def deal_utterances(utts, cores):
    """Longest utterance first onto the least loaded core: every core gets about the same seconds of audio, so
    the wall time of a pass is the machine's throughput and not its unluckiest core (round 1 dealt round-robin)."""
    from hts_train_world_b200 import signals
    dur = {u: signals.utterance_params(u)["T"] for u in utts}
    bins, load = [[] for _ in range(cores)], [0.0] * cores
    for u in sorted(utts, key=lambda v: (-dur[v], v)):
        i = min(range(cores), key=lambda k: (load[k], k))
        bins[i].append(u)
        load[i] += dur[u]
    return bins


class ReferencePool:
    """`cores` processes; utterances are dealt round-robin."""

    def __init__(self, cores, opt=True):
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        self.cores = cores
        self.conns, self.procs = [], []
        for _ in range(cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker_main, args=(b, opt), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)

    def _all(self, cmd, utts):
        for cn, mine in zip(self.conns, deal_utterances(utts, self.cores)):
            cn.send((cmd, mine))
        return [cn.recv() for cn in self.conns]

    def prepare(self, utts):
        self._all("prep", utts)

    def run(self, utts):
        t0 = time.perf_counter()
        stages = self._all("run", utts)
        dt = time.perf_counter() - t0
        return dt, np.sum(np.asarray(stages), axis=0)

    def close(self):
        for cn in self.conns:
            cn.send(("quit", None))
        for p in self.procs:
            p.join(timeout=10)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_sample_utts(cores, per_core):
    return list(range(cores * per_core))


def run_reference_pass(pool, utts, passes):
    """-> (audio seconds per pass, [seconds per pass], per-stage cpu seconds summed)."""
    from hts_train_world_b200 import signals
    audio = sum(signals.utterance_params(u)["T"] for u in utts)
    pool.prepare(utts)          # synthetic signals are generated untimed, inside each worker
    times, stages = [], np.zeros(6)
    for _ in range(passes):
        dt, st = pool.run(utts)
        times.append(dt)
        stages += st
    return audio, times, stages


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    opt = True
    if not os.path.exists(ref.ref_path(opt)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libworld_ref_O3.so was not built "
                          "(needs /root/reference at build time)"}))
        return 0
    cores = host_cores()
    utts = reference_sample_utts(cores, 4)
    pool = ReferencePool(cores, opt)
    audio, times, stages = run_reference_pass(pool, utts, args.warmup + args.steps)
    pool.close()
    timed = times[args.warmup:]
    total = float(sum(timed))
    value = audio * len(timed) / total
    sample = ("%d utterances (ids 0..%d, %.1f s of 48 kHz audio) per step, one process per core, utterances dealt "
              "longest first onto the least loaded core; reference WORLD_v2 sources compiled -O3 (no -ffast-math)" % (len(utts), len(utts) - 1, audio))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "1132-utterance synthetic 48 kHz corpus, 5 ms frames, fft_size 2048: "
                               "Dio+StoneMask+CheapTrick+D4C+codec+Synthesis (bounded sample per step)",
                   "fs": FS, "frame_period_ms": FRAME_PERIOD, "utterances_per_step": len(utts)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_cpu_seconds": dict(zip(["dio", "stonemask", "cheaptrick", "d4c", "synthesis", "codec"],
                                      [float(s) for s in stages])),
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# =================================================================================================
# our arm
# =================================================================================================
class ClockSampler:
    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [v.strip() for v in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def pin_rank_to_cores(local, world):
    """One process per GPU: give every rank its own slice of the host cores (contiguous, so that the ranks of
    the GPUs of one socket stay on that socket) and keep torch's intra-op pool small.  Eight unpinned ranks
    wander over all cores and their launch / read-back threads contend (SCALE_r01: end-to-end efficiency
    0.935 at N = 8 with device-timed efficiency 0.995)."""
    try:
        import torch
        torch.set_num_threads(max(1, min(4, host_cores() // max(1, world))))
        if world <= 1:
            return
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per >= 2:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
    except Exception:
        pass


def build_shard(args, rank, world):
    """Utterance ids of this rank.  Weak scaling: the job is world * utts utterances, dealt
    longest-first to the ranks (hts-train-world_b200/corpus.py)."""
    from hts_train_world_b200 import corpus, signals
    n_total = args.utts * world
    lengths = [int(round(signals.utterance_params(u)["T"] * FS)) for u in range(n_total)]
    ids = corpus.shard_utterances(lengths, rank, world)
    return ids, [lengths[i] for i in ids]


def verify_against_reference(wb, c, ids, pcm_host, lengths, k, f0_method="dio"):
    """-> the five north_star metrics over k utterances of the resident batch `c` (worst case of each)."""
    from oracle import metrics as M
    from oracle import ref
    if not os.path.exists(ref.ref_path()):
        return {"unavailable": "oracle/_ref/libworld_ref.so was not built"}
    R = ref.load()
    n = len(lengths)
    pick = sorted(set(int(round(i * (n - 1) / max(1, k - 1))) for i in range(k))) if n > 1 else [0]
    offs = np.concatenate([[0], np.cumsum(lengths)])
    res = {"vuv_agreement": 1.0, "f0_rel_error": 0.0, "lsd_db_max": 0.0, "ap_abs_error": 0.0, "snr_db": 1e9}
    fft = c.fft_size
    for u in pick:
        x = pcm_host[int(offs[u]):int(offs[u + 1])].numpy().astype(np.float64) / 32768.0
        o = c.utterance(u)
        if f0_method == "harvest":
            tp, f0_ref = R.harvest(x, FS, frame_period=FRAME_PERIOD)
        else:
            tp, f0r = R.dio(x, FS, frame_period=FRAME_PERIOD)
            f0_ref = R.stonemask(x, FS, tp, f0r)
        res["vuv_agreement"] = min(res["vuv_agreement"], M.vuv_agreement(f0_ref, o["f0"]))
        res["f0_rel_error"] = max(res["f0_rel_error"], M.f0_rel_error(f0_ref, o["f0"]))
        sp_ref = R.cheaptrick(x, FS, tp, o["f0"])
        ap_ref = R.d4c(x, FS, tp, o["f0"], fft, threshold=0.0)
        res["lsd_db_max"] = max(res["lsd_db_max"], M.lsd_db(sp_ref, o["sp"])[1])
        res["ap_abs_error"] = max(res["ap_abs_error"], M.ap_abs_error(ap_ref, o["ap"]))
        y_ref = R.synthesis(o["f0"], o["sp"], o["ap"], fft, FRAME_PERIOD, FS)
        res["snr_db"] = min(res["snr_db"], M.snr_db(y_ref, o["y"]))
    res = {a: float(b) for a, b in res.items()}
    res["within_tolerance"] = bool(res["vuv_agreement"] >= M.TOL_VUV_AGREEMENT and res["f0_rel_error"] <= M.TOL_F0_REL and
                                   res["lsd_db_max"] <= M.TOL_LSD_DB and res["ap_abs_error"] <= M.TOL_AP_ABS and
                                   res["snr_db"] >= M.TOL_SNR_DB)
    res["tolerances"] = {"vuv_agreement": M.TOL_VUV_AGREEMENT, "f0_rel_error": M.TOL_F0_REL, "lsd_db_max": M.TOL_LSD_DB,
                         "ap_abs_error": M.TOL_AP_ABS, "snr_db": M.TOL_SNR_DB}
    res["utterances"] = [int(ids[u]) for u in pick]
    res["how"] = ("f0 / sp / ap / y of these utterances read back from the batch the timed region processed; reference = "
                  "oracle/_ref (unmodified WORLD_v2); CheapTrick, D4C and Synthesis compared stage by stage on identical inputs")
    return res


def other_configs(wb, c, args, audio_s, pcm_dev, lengths, stream):
    """BASELINE.json configs[0], [2], [3] measured in the same run (configs[1] is the headline line,
    configs[4] is this bench under torchrun).  Device-timed with CUDA events on the library stream."""
    import torch
    out = {}

    def dev_timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    # configs[3]: Synthesis only, f0 / sp / ap of the batch already in HBM; the batch is resynthesised
    # ceil(10 h / batch) times back to back inside one timed region
    passes = max(1, int(math.ceil(36000.0 / audio_s)))
    t0 = time.perf_counter()
    ms = dev_timed(lambda: [c.synthesis() for _ in range(passes)], 1, 1)
    out["config4_synthesis_only"] = {
        "workload": "Synthesis only: %d passes over the %d-utterance batch = %.2f h of 48 kHz audio, f0/sp/ap (double) resident in HBM"
                    % (passes, len(lengths), passes * audio_s / 3600.0),
        "value": passes * audio_s / (ms * 1e-3), "unit": UNIT, "ms_total": ms, "wall_s": time.perf_counter() - t0}
    # configs[2]: the Harvest F0 path (71-800 Hz) instead of Dio + StoneMask on the same corpus
    if args.f0 != "harvest":
        def step_h():
            c.set_pcm16_device(pcm_dev)
            c.analyze(f0="harvest")
            c.code(MGC_DIM, BAP_DIM)
            c.synthesis()
            c.feature_stats()
        try:
            wb.kernel_timing(True)
            wb.kernel_times_reset()
            ms = dev_timed(step_h, 2, 3)
            hk = {k: wb.kernel_time(k) for k in ["harvest_iir_kernel", "harvest_filter_kernel", "harvest_zc_kernel", "harvest_raw_kernel",
                                                 "harvest_refine_kernel", "harvest_unreliable_kernel", "harvest_fix_a_kernel",
                                                 "harvest_fix_kernel", "harvest_smooth_kernel"]}
            wb.kernel_timing(False)
            out["config3_harvest"] = {
                "workload": "%d-utterance corpus, Harvest (71-800 Hz) + CheapTrick + D4C + codec + Synthesis + statistics" % len(lengths),
                "value": audio_s / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "harvest_stage_ms": wb.stage_times()["harvest"],
                "kernels_ms_per_step": {k: v[0] / 5.0 for k, v in hk.items() if v[1]}}
            if args.verify > 0:       # F0 of the first and the last utterance of the batch against the reference's Harvest
                from oracle import metrics as M
                from oracle import ref
                if os.path.exists(ref.ref_path()):
                    R = ref.load()
                    offs = np.concatenate([[0], np.cumsum(lengths)])
                    agree, err = 1.0, 0.0
                    for u in (0, len(lengths) - 1):
                        x = pcm_dev[int(offs[u]):int(offs[u + 1])].cpu().numpy().astype(np.float64) / 32768.0
                        _, f0_ref = R.harvest(x, FS, frame_period=FRAME_PERIOD)
                        f0_u = c.utterance(u)["f0"]
                        agree = min(agree, M.vuv_agreement(f0_ref, f0_u))
                        err = max(err, M.f0_rel_error(f0_ref, f0_u))
                    out["config3_harvest"]["parity"] = {"vuv_agreement": agree, "f0_rel_error": err,
                                                        "within_tolerance": bool(agree >= M.TOL_VUV_AGREEMENT and err <= M.TOL_F0_REL)}
        except Exception as e:          # e.g. out of memory for the band signals of a very large batch
            out["config3_harvest"] = {"error": str(e)[:300]}
    # configs[0]: one 3 s 16 kHz utterance through the drop-in C API, host buffers, one call per stage
    try:
        from hts_train_world_b200 import signals
        x = signals.pcm_to_double(signals.make_utterance(11, 16000, duration=3.0)[0])

        def chain():
            tp, f0r = wb.dio(x, 16000)
            f0 = wb.stonemask(x, 16000, tp, f0r)
            sp = wb.cheaptrick(x, 16000, tp, f0)
            ap = wb.d4c(x, 16000, tp, f0, 1024, threshold=0.0)
            wb.synthesis(f0, sp, ap, 1024, 5.0, 16000)
        for _ in range(3):
            chain()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            chain()
            ts.append(time.perf_counter() - t0)
        out["config1_dropin"] = {"workload": "single 3 s synthetic 16 kHz utterance, drop-in C API, host buffers, one call per stage",
                                 "value": 3.0 / min(ts), "unit": UNIT, "ms_total": 1e3 * min(ts)}
    except Exception as e:
        out["config1_dropin"] = {"error": str(e)[:300]}
    return out



def ours_arm(args):
    import torch
    import torch.distributed as dist
    import hts_train_world_b200 as wb
    from hts_train_world_b200 import corpus, roofline, signals

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU reference)")
    cpu_pool = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref
        if os.path.exists(ref.ref_path(True)):
            cpu_pool = ReferencePool(host_cores(), True)    # forked before CUDA is initialised; idle until the end
    torch.cuda.set_device(local)
    pin_rank_to_cores(local, world)
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective; stdout carries ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    wb.init(local)
    stream = torch.cuda.Stream()
    wb.set_stream(stream.cuda_stream)

    # ---- synthetic corpus shard: generated on the GPU, then parked in pinned host memory ---------
    ids, lengths = build_shard(args, rank, world)
    t_gen = time.perf_counter()
    pcm_dev = torch.empty(sum(lengths), dtype=torch.int16, device="cuda")
    o = 0
    for u, n in zip(ids, lengths):
        p = signals.make_utterance(int(u), FS, device="cuda")[0]
        assert p.numel() == n
        pcm_dev[o:o + n] = p
        o += n
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    pcm_host = torch.empty_like(pcm_dev, device="cpu").pin_memory()
    pcm_host.copy_(pcm_dev)
    audio_s = sum(lengths) / float(FS)

    c = wb.Corpus(FS, lengths, FRAME_PERIOD)
    f0_host = torch.empty(c.total_frames, dtype=torch.float64).pin_memory()
    lf0_host = torch.empty(c.total_frames, dtype=torch.float32).pin_memory()
    mgc_host = torch.empty((c.total_frames, MGC_DIM), dtype=torch.float32).pin_memory()
    bap_host = torch.empty((c.total_frames, BAP_DIM), dtype=torch.float32).pin_memory()

    def reduce_stats(st):
        if world > 1:
            t = torch.as_tensor(st, dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            return t.cpu().numpy()
        return st

    def step_resident():
        c.set_pcm16_device(pcm_dev)
        c.analyze(f0=args.f0)
        c.code(MGC_DIM, BAP_DIM)
        c.synthesis()
        return reduce_stats(c.feature_stats())

    # ---- end to end: the shard goes through the public batch API in E2E_PARTS pipelined sub-batches.
    # Uploads run on the library's upload stream one sub-batch ahead, the coded features and the 16-bit
    # waveform leave on its download stream while the next sub-batch computes.
    n_parts = max(1, min(E2E_PARTS, len(lengths)))
    bounds = [len(lengths) * i // n_parts for i in range(n_parts + 1)]
    parts = []
    s_off = f_off_ = y_off_ = 0
    for i in range(n_parts):
        ls = lengths[bounds[i]:bounds[i + 1]]
        cp = wb.Corpus(FS, ls, FRAME_PERIOD)
        ns, nf = sum(ls), cp.total_frames
        ny = sum(int((int(f) - 1) * FRAME_PERIOD / 1000.0 * FS) + 1 for f in cp.f_len)     # W/test/synth.cpp:259
        parts.append(dict(c=cp, s=(s_off, s_off + ns), f=(f_off_, f_off_ + nf), y=(y_off_, y_off_ + ny)))
        s_off, f_off_, y_off_ = s_off + ns, f_off_ + nf, y_off_ + ny
    y_host = torch.empty(y_off_, dtype=torch.int16).pin_memory()
    # Steps follow each other like the batches of a long corpus run: the first sub-batch of the next
    # step is uploaded while the last one of this step computes, and the results of a step land in one
    # of two sets of pinned host buffers while the next step already runs (a consumer has one step of
    # time to take them).  Nothing is skipped: every step uploads its PCM and downloads all its
    # results inside the timed region; the statistics are read synchronously every sub-batch.
    out_sets = [dict(f0=f0_host, lf0=lf0_host, mgc=mgc_host, bap=bap_host, y=y_host)]
    out_sets.append({k: torch.empty_like(v).pin_memory() for k, v in out_sets[0].items()})
    e2e_state = dict(step=0, prefetched=False)

    e2e_skip = set(filter(None, os.environ.get("WB_E2E_SKIP", "").split(",")))   # developer aid: leave copies out to see what each costs

    def upload_part(q):
        if "up" in e2e_skip and e2e_state["step"] > 1:
            return
        q["c"].upload_pcm16_async(pcm_host[q["s"][0]:q["s"][1]])

    def step_e2e(more_to_come=True):
        st = np.zeros((1 + MGC_DIM, 3))
        o = out_sets[e2e_state["step"] % 2]
        e2e_state["step"] += 1
        if not e2e_state["prefetched"]:
            upload_part(parts[0])
        e2e_state["prefetched"] = False
        for i, p in enumerate(parts):
            if i + 1 < len(parts):
                upload_part(parts[i + 1])
            elif more_to_come:
                upload_part(parts[0])               # first sub-batch of the next step
                e2e_state["prefetched"] = True
            cp, (fa, fb), (ya, yb) = p["c"], p["f"], p["y"]
            # The synchronous reads (f0, statistics) come BEFORE the bulk asynchronous copies are
            # queued: a read issued behind a bulk copy waits in the device-to-host copy engine's
            # queue for that copy's whole PCIe time, and the host cannot launch the next stage
            # meanwhile (that was the 15 ms per step the end-to-end leg lost, profiles/README.md).
            cp.analyze(f0=args.f0)
            if "f0" not in e2e_skip:
                wb._check(wb.lib().wb200_batch_get_f0(cp._h, wb.C.cast(o["f0"][fa:fb].data_ptr(), wb._dp), 1), "get_f0")
            cp.code(MGC_DIM, BAP_DIM)
            st += cp.feature_stats()
            # Both bulk copies are queued BEHIND Synthesis: its first milliseconds read three small counts back
            # (pulse bounds, list sizes), and a read-back completes only when a bulk device-to-host copy that is
            # in flight has drained -- the features' copy (114 MB per sub-batch) cost its whole PCIe time when it
            # was queued in front of Synthesis (measured: 198.9 -> 194.5 ms per step).  Behind Synthesis the copies
            # run beside the next sub-batch's Dio, whose first read-back comes after its 5 ms filter kernel.
            early = "early" in e2e_skip
            if "coded" not in e2e_skip and early:
                cp.coded_async(o["lf0"][fa:fb], o["mgc"][fa:fb], o["bap"][fa:fb])
            cp.synthesis()
            if "wave" not in e2e_skip:
                cp.y_pcm16_async(o["y"][ya:yb])
            if "coded" not in e2e_skip and not early:
                cp.coded_async(o["lf0"][fa:fb], o["mgc"][fa:fb], o["bap"][fa:fb])
        if not more_to_come:
            wb.sync()                           # every asynchronous copy has landed
        return reduce_stats(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, (t1 - t0) * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        else:
            wall = (t1 - t0) * 1e3
        return ms, wall, out, (t0, t1)

    for _ in range(args.warmup):
        step_resident()
    wb.kernel_timing(True)
    wb.kernel_times_reset()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    n0 = wb.launch_count()
    ms, wall, stats, (t0, t1) = timed(step_resident, args.steps)
    launches = wb.launch_count() - n0
    kernel_ms = {k: wb.kernel_time(k) for k in
                 ["d4c_main_kernel", "d4c_gd_kernel", "d4c_tail_kernel", "d4c_lovetrain_kernel", "cheaptrick_kernel", "synth_pulse_kernel",
                  "synth_timebase_kernel", "stonemask_kernel", "dio_filter_kernel", "dio_zc_kernel",
                  "dio_candidates_kernel", "dio_fix_kernel", "harvest_iir_kernel", "harvest_filter_kernel",
                  "harvest_zc_kernel", "harvest_raw_kernel", "harvest_refine_kernel", "harvest_unreliable_kernel",
                  "harvest_fix_a_kernel", "harvest_fix_kernel", "harvest_smooth_kernel",
                  "codec_encode_kernel"]}
    wb.kernel_timing(False)
    stage_ms = wb.stage_times()
    clk = clocks.stop(t0, t1) if rank == 0 else None
    if world > 1:
        ta = torch.tensor([audio_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(ta)
        audio_total = float(ta[0])
    else:
        audio_total = audio_s
    value = audio_total * args.steps / (ms * 1e-3)

    if os.environ.get("WB_STEP_TRACE"):       # developer aid: wall time of every call of one resident step (stream drained in between)
        def traced(name, fn):
            torch.cuda.synchronize()
            t_a = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            sys.stderr.write("[step] %-18s %8.2f ms\n" % (name, 1e3 * (time.perf_counter() - t_a)))
            return r
        for _ in range(2):
            traced("set_pcm16_device", lambda: c.set_pcm16_device(pcm_dev))
            if args.f0 == "harvest":
                traced("harvest", c.harvest)
            else:
                traced("dio", c.dio)
                traced("stonemask", c.stonemask)
            traced("cheaptrick", c.cheaptrick)
            traced("d4c", lambda: c.d4c(threshold=0.0))
            traced("code", lambda: c.code(MGC_DIM, BAP_DIM))
            traced("synthesis", c.synthesis)
            traced("feature_stats", c.feature_stats)

    # ---- end to end from host memory --------------------------------------------------------------
    # deferred copies: the bulk uploads / downloads of the pipeline start right before D4C's main kernel instead
    # of beside a stage whose read-backs they would delay (include/world_b200.h); WB_E2E_DEFER=0 switches it off
    defer = os.environ.get("WB_E2E_DEFER", "1") != "0"
    wb.set_copy_deferral(defer)
    step_e2e()
    step_e2e(more_to_come=False)
    e2e_calls = dict(n=0)

    def step_e2e_timed():
        e2e_calls["n"] += 1
        return step_e2e(more_to_come=e2e_calls["n"] < args.steps)

    if os.environ.get("WB_E2E_TRACE"):        # developer aid: wall time of every call of one end-to-end step, device drained in between
        def traced_e(name, fn):
            torch.cuda.synchronize()
            t_a = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            sys.stderr.write("[e2e step] %-16s %8.2f ms\n" % (name, 1e3 * (time.perf_counter() - t_a)))
            return r
        o = out_sets[0]
        for i, q in enumerate(parts):
            cp, (fa, fb), (ya, yb) = q["c"], q["f"], q["y"]
            traced_e("upload", lambda: upload_part(q))
            traced_e("dio", cp.dio)
            traced_e("stonemask", cp.stonemask)
            traced_e("cheaptrick", cp.cheaptrick)
            traced_e("d4c", lambda: cp.d4c(threshold=0.0))
            traced_e("get_f0", lambda: wb._check(wb.lib().wb200_batch_get_f0(cp._h, wb.C.cast(o["f0"][fa:fb].data_ptr(), wb._dp), 1), "get_f0"))
            traced_e("code", lambda: cp.code(MGC_DIM, BAP_DIM))
            traced_e("feature_stats", cp.feature_stats)
            traced_e("coded_async", lambda: cp.coded_async(o["lf0"][fa:fb], o["mgc"][fa:fb], o["bap"][fa:fb]))
            traced_e("synthesis", cp.synthesis)
            traced_e("y_pcm16_async", lambda: cp.y_pcm16_async(o["y"][ya:yb]))
        wb.sync()
        e2e_state["prefetched"] = False
    e2e_ktime = None
    if os.environ.get("WB_E2E_KTIME"):        # developer aid: the library's per-kernel event timers during the end-to-end leg
        wb.kernel_timing(True)
        wb.kernel_times_reset()
    ms_e, wall_e, _, _ = timed(step_e2e_timed, args.steps)
    if os.environ.get("WB_E2E_KTIME"):
        e2e_ktime = {k: wb.kernel_time(k)[0] / args.steps for k in kernel_ms}
        e2e_ktime = {k: v for k, v in e2e_ktime.items() if v}
        e2e_ktime["sum"] = sum(e2e_ktime.values())
        e2e_ktime["stages_last_sub_batch"] = wb.stage_times()
        wb.kernel_timing(False)
        sys.stderr.write("[e2e] rank %d kernels %s\n" % (rank, json.dumps(e2e_ktime)))
    wb.set_copy_deferral(False)
    e2e_value = audio_total * args.steps / (max(ms_e, wall_e) * 1e-3)
    h2d = pcm_host.numel() * 2
    d2h = y_host.numel() * 2 + f0_host.numel() * 8 + (lf0_host.numel() + mgc_host.numel() + bap_host.numel()) * 4 + \
        (1 + MGC_DIM) * 24

    # ---- parity of the timed configuration (--verify K) ------------------------------------------------
    # K utterances of the batch the timed region just processed are pulled out of HBM (f0, sp, ap of the
    # resident leg's last step, y of its Synthesis) and compared with the compiled reference: F0 against
    # the reference's own Dio + StoneMask chain; CheapTrick / D4C / Synthesis stage by stage, the reference
    # fed the same upstream values our stage saw -- north_star's five metrics.
    parity = None
    if rank == 0 and args.verify > 0:
        parity = verify_against_reference(wb, c, ids, pcm_host, lengths, args.verify, args.f0)

    # ---- the other BASELINE.json configurations (rank 0, N = 1) -------------------------------------------
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        configs = other_configs(wb, c, args, audio_s, pcm_dev, lengths, stream)

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    f0 = f0_host.numpy().copy()
    voiced = f0 > 0
    pv = float((f0[voiced] * FRAME_PERIOD / 1000.0).sum())
    pu = float((~voiced).sum() * 500.0 * FRAME_PERIOD / 1000.0)
    counts = roofline.stage_counts(FS, f0, sum(lengths), pv, pu, fft_size=c.fft_size, n_utt=len(lengths),
                                   lovetrain_fp32=wb.build_info().get("lovetrain_fp32", False))
    fp64_peak = wb.fma_peak_tflops(True)
    fp32_peak = wb.fma_peak_tflops(False)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    kmap = {"d4c_main_kernel": "d4c_main", "d4c_gd_kernel": "d4c_gd", "d4c_tail_kernel": "d4c_tail",
            "d4c_lovetrain_kernel": "d4c_lovetrain", "cheaptrick_kernel": "cheaptrick",
            "synth_pulse_kernel": "synthesis", "stonemask_kernel": "stonemask", "dio_filter_kernel": "dio",
            "codec_encode_kernel": "codec"}
    # Every kernel against the peak of the precision it EXECUTES in (roofline.py splits the algorithmic
    # FLOPs of each kernel into its FP64 and FP32 part): bound time = flops64 / peak64 + flops32 / peak32,
    # or the compulsory bytes over the HBM bandwidth when that is longer.
    kernels = {}
    for k, (tot_ms, n) in kernel_ms.items():
        if n == 0:
            continue
        ent = {"ms_per_launch": tot_ms / n, "launches_per_step": n / args.steps}
        if k in kmap and kmap[k] in counts:
            cnt = counts[kmap[k]]
            per_step_s = tot_ms * 1e-3 / args.steps        # all launches of this kernel in one step
            t_roof, bound = roofline.roof_seconds(cnt, fp64_peak, fp32_peak, hbm_peak)
            ent.update({"tflops": cnt["flops"] / per_step_s / 1e12, "fp32_share_of_flops": cnt.get("flops32", 0.0) / cnt["flops"],
                        "gbs": cnt["bytes"] / per_step_s / 1e9, "bound": bound, "frac": t_roof / per_step_s,
                        "frac_hbm": cnt["bytes"] / per_step_s / 1e9 / hbm_peak})
        kernels[k] = ent
    dom = max((k for k in kernels if k in kmap), key=lambda k: kernel_ms[k][0])
    dcnt = counts[kmap[dom]]
    dsec = kernel_ms[dom][0] * 1e-3 / args.steps
    t_roof, bound = roofline.roof_seconds(dcnt, fp64_peak, fp32_peak, hbm_peak)
    if bound == "hbm":
        roof = {"bound": "hbm", "achieved": dcnt["bytes"] / dsec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}
    else:
        # `peak` is the rate at which THIS kernel's precision mix would run with every pipe it uses at its
        # measured FMA peak: flops / (flops64 / peak64 + flops32 / peak32); an all-FP64 kernel gets peak64.
        roof = {"bound": bound, "achieved": dcnt["flops"] / dsec / 1e12, "peak": dcnt["flops"] / t_roof / 1e12,
                "unit": "TFLOP/s",
                "peak_source": "measured live on this GPU: FMA micro-benchmarks, FP64 %.2f and FP32 %.2f TFLOP/s (nominal 37.2 / 74.4), "
                               "weighted by the kernel's FP64 / FP32 split of algorithmic FLOPs" % (fp64_peak, fp32_peak)}
    roof.update({"frac": roof["achieved"] / roof["peak"], "traffic": None, "kernel": dom,
                 "algorithmic_flops_per_launch": dcnt["flops"], "algorithmic_fp32_flops_per_launch": dcnt.get("flops32", 0.0),
                 "algorithmic_bytes_per_launch": dcnt["bytes"],
                 "units_per_launch": dcnt["units"], "ms_per_launch": dsec * 1e3,
                 "share_of_step": kernel_ms[dom][0] / ms,
                 "note": "every stage of this path is bound by the CUDA-core floating-point pipes, not HBM or tensor "
                         "cores (SURVEY.md 8d); hbm fraction of the same kernel: %.4f" %
                         (dcnt["bytes"] / dsec / 1e9 / hbm_peak)})
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            roof["traffic"] = json.load(open(tp)).get(dom)
        except Exception:
            pass
    # the whole step: sum of the algorithmic work of every stage over the step time
    step_keys = ["dio", "stonemask", "cheaptrick", "d4c", "codec", "synthesis"] if args.f0 != "harvest" else \
        ["cheaptrick", "d4c", "codec", "synthesis"]
    step_cnt = {"flops": sum(counts[k]["flops"] for k in step_keys), "flops32": sum(counts[k].get("flops32", 0.0) for k in step_keys),
                "bytes": sum(counts[k]["bytes"] for k in step_keys)}
    t_step_roof, step_bound = roofline.roof_seconds(step_cnt, fp64_peak, fp32_peak, hbm_peak)
    step_s = ms * 1e-3 / args.steps
    roofline_step = {"bound": step_bound, "achieved": step_cnt["flops"] / step_s / 1e12, "peak": step_cnt["flops"] / t_step_roof / 1e12,
                     "unit": "TFLOP/s", "frac": t_step_roof / step_s, "algorithmic_flops": step_cnt["flops"],
                     "algorithmic_fp32_flops": step_cnt["flops32"], "algorithmic_bytes": step_cnt["bytes"],
                     "stages": step_keys, "flops_per_audio_s": step_cnt["flops"] / audio_s}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%d-utterance synthetic 48 kHz corpus per GPU (%.0f s audio, %d frames), 5 ms "
                               "frames, fft_size %d: %s+CheapTrick+D4C+codec+Synthesis + lf0/mgc statistics"
                               % (len(lengths), audio_s, c.total_frames, c.fft_size,
                                  "Harvest" if args.f0 == "harvest" else "Dio+StoneMask"),
                   "fs": FS, "frame_period_ms": FRAME_PERIOD, "utterances_per_gpu": len(lengths),
                   "parallelism": "utterance-sharded x%d, no hot-path collective" % world,
                   "l2": "inputs larger than L2 (%.0f MB PCM, GBs of intermediates per step)" % (h2d / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": max(ms_e, wall_e) / args.steps, "pipelined_sub_batches": n_parts, "deferred_copies": defer,
                "pipelined_across_steps": "the next step's first upload overlaps this step's last sub-batch; results land in double-buffered pinned host memory",
                "device_ms_per_step": ms_e / args.steps, "wall_ms_per_step": wall_e / args.steps,
                "result": "the analysis tool's float32 lf0/mgc/bap + f0 + 16-bit resynthesised waveform + lf0/mgc statistics; "
                          "sp/ap stay in HBM"},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall / args.steps,
        "stage_ms": stage_ms,
        "roofline": roof,
        "roofline_step": roofline_step,
        # the same kernel against the memory roofline (it is nowhere near it: the frame lives in shared memory)
        "roofline_hbm": {"bound": "hbm", "achieved": dcnt["bytes"] / dsec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": dcnt["bytes"] / dsec / 1e9 / hbm_peak, "traffic": roof.get("traffic"), "kernel": dom},
        "parity": parity,
        "configs": configs,
        "dtype_note": "FP64 wherever a cancellation follows (power spectra, cumulative sums, group delay, time base); "
                      "FP32 transforms for log spectra / cepstra / noise / band slices (DESIGN.md section 4)",
        "kernels": kernels,
        "peaks": {"fp64_tflops_measured": fp64_peak, "fp32_tflops_measured": fp32_peak, "hbm_gbs": hbm_peak},
        "clocks": clk,
        "lf0_stats": corpus.merge_stats([stats[0]]),
        "mgc0_stats": corpus.merge_stats([stats[1]]),
        "gen_seconds": t_gen,
    }

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if cpu_pool is not None:
            cores = cpu_pool.cores
            utts = reference_sample_utts(cores, 4)
            audio, times, stages = run_reference_pass(cpu_pool, utts, 1)
            cpu_pool.close()
            line["cpu_baseline"] = {
                "value": audio / times[0], "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": "%d utterances (ids 0..%d, %.1f s audio), one pass, one process per core (longest first onto the least loaded "
                          "core), reference sources compiled -O3" % (len(utts), len(utts) - 1, audio),
                "stage_cpu_seconds": dict(zip(["dio", "stonemask", "cheaptrick", "d4c", "synthesis", "codec"],
                                              [float(s) for s in stages]))}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": "oracle/_ref not built"}
    if rank == 0:
        print(json.dumps(line))
    c.close()
    for p in parts:
        p["c"].close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=1132, help="utterances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", type=int, default=3, help="utterances of the timed batch checked against the compiled reference")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs of BASELINE configs 1, 3 and 4")
    ap.add_argument("--f0", default="dio", choices=["dio", "harvest"], help="F0 estimator (harvest = BASELINE config 3)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours_arm(args)


if __name__ == "__main__":
    sys.exit(main())

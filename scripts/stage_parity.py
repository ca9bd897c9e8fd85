"""Developer tool (GPU box): per-stage parity of libworld_b200.so against the compiled
reference (oracle/_ref), each stage fed the ORACLE's upstream outputs so errors do not
compound.  Usage: python scripts/stage_parity.py [stages...] [--fs 48000] [--utts 3]"""
import argparse
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hts_train_world_b200 as wb
from hts_train_world_b200 import signals
from oracle import metrics as M
from oracle import ref

ap = argparse.ArgumentParser()
ap.add_argument("stages", nargs="*", default=["rng", "cheaptrick", "d4c", "stonemask", "dio", "synthesis"])
ap.add_argument("--fs", type=int, default=48000)
ap.add_argument("--utts", type=int, default=2)
ap.add_argument("--dur", type=float, default=2.0)
args = ap.parse_args()

R = ref.load()
wb.init(0)
ok_all = True


def report(name, ok, msg):
    global ok_all
    ok_all &= bool(ok)
    print("%-12s %s  %s" % (name, "PASS" if ok else "FAIL", msg), flush=True)


if "rng" in args.stages:
    n = 20000
    a, b = ref.randn_stream(n), wb.randn_stream(n)
    report("rng", np.array_equal(a, b), "first %d variates bit-exact=%s" % (n, np.array_equal(a, b)))

for u in range(args.utts):
    pcm, _ = signals.make_utterance(u, args.fs, duration=args.dur)
    x = signals.pcm_to_double(pcm)
    fs = args.fs
    o = R.analyze(x, fs)
    t, f0, fft = o["t"], o["f0"], o["fft_size"]
    for st in args.stages:
        try:
            t0 = time.time()
            if st == "cheaptrick":
                sp = wb.cheaptrick(x, fs, t, f0)
                mean, mx = M.lsd_db(o["sp"], sp)
                report("cheaptrick", mx <= M.TOL_LSD_DB, "utt %d LSD mean %.3e max %.3e dB (%.2fs)" % (u, mean, mx, time.time() - t0))
            elif st == "d4c":
                apn = wb.d4c(x, fs, t, f0, fft)
                e = M.ap_abs_error(o["ap"], apn)
                report("d4c", e <= M.TOL_AP_ABS, "utt %d ap max abs err %.3e (%.2fs)" % (u, e, time.time() - t0))
            elif st == "stonemask":
                f = wb.stonemask(x, fs, t, o["f0_raw"])
                report("stonemask", M.vuv_agreement(f0, f) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(f0, f) <= M.TOL_F0_REL,
                       "utt %d vuv %.5f f0 rel %.3e" % (u, M.vuv_agreement(f0, f), M.f0_rel_error(f0, f)))
            elif st == "dio":
                tt, f = wb.dio(x, fs)
                report("dio", M.vuv_agreement(o["f0_raw"], f) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(o["f0_raw"], f) <= M.TOL_F0_REL,
                       "utt %d vuv %.5f f0 rel %.3e t_eq %s" % (u, M.vuv_agreement(o["f0_raw"], f), M.f0_rel_error(o["f0_raw"], f), np.array_equal(tt, t)))
            elif st == "synthesis":
                yr = R.synthesis(f0, o["sp"], o["ap"], fft, 5.0, fs)
                yn = wb.synthesis(f0, o["sp"], o["ap"], fft, 5.0, fs)
                s = M.snr_db(yr, yn)
                report("synthesis", s >= M.TOL_SNR_DB, "utt %d SNR %.1f dB" % (u, s))
            elif st == "harvest":
                _, fr = R.harvest(x, fs)
                _, f = wb.harvest(x, fs)
                report("harvest", M.vuv_agreement(fr, f) >= M.TOL_VUV_AGREEMENT and M.f0_rel_error(fr, f) <= M.TOL_F0_REL,
                       "utt %d vuv %.5f f0 rel %.3e" % (u, M.vuv_agreement(fr, f), M.f0_rel_error(fr, f)))
        except Exception as e:
            traceback.print_exc()
            report(st, False, "exception: %s" % e)
print("launches:", wb.launch_count(), "stage ms:", wb.stage_times())
sys.exit(0 if ok_all else 1)

"""BASELINE.json configs[0]: one 3 s synthetic 16 kHz utterance, 5 ms frames, through the DROP-IN
WORLD C API (Dio, StoneMask, CheapTrick, D4C, Synthesis called one after the other exactly as
W/test/analysis.cpp + synth.cpp do), timed beside the compiled reference on one host core.
Usage (GPU box): python scripts/config1_latency.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hts_train_world_b200 as wb
from hts_train_world_b200 import signals
from oracle import metrics as M
from oracle import ref

fs = 16000
x = signals.pcm_to_double(signals.make_utterance(11, fs, duration=3.0)[0])
R = ref.load(opt=True)
wb.init(0)


def chain(L):
    t = [time.perf_counter()]
    tp, f0r = L.dio(x, fs); t.append(time.perf_counter())
    f0 = L.stonemask(x, fs, tp, f0r); t.append(time.perf_counter())
    fft = 1024
    sp = L.cheaptrick(x, fs, tp, f0); t.append(time.perf_counter())
    ap = L.d4c(x, fs, tp, f0, fft, threshold=0.0); t.append(time.perf_counter())
    y = L.synthesis(f0, sp, ap, fft, 5.0, fs); t.append(time.perf_counter())
    return np.diff(t) * 1e3, (f0, sp, ap, y)


for _ in range(3):
    chain(wb)
ours = np.min([chain(wb)[0] for _ in range(10)], axis=0)
refms = np.min([chain(R)[0] for _ in range(3)], axis=0)
o, r = chain(wb)[1], chain(R)[1]
res = {"config": "single 3 s synthetic 16 kHz utterance, 5 ms frames, drop-in C API (one call per stage, host buffers)",
       "stages": ["dio", "stonemask", "cheaptrick", "d4c", "synthesis"],
       "ours_ms": [round(float(v), 3) for v in ours], "ours_total_ms": round(float(ours.sum()), 3),
       "reference_1core_O3_ms": [round(float(v), 3) for v in refms], "reference_total_ms": round(float(refms.sum()), 3),
       "xRT_ours": round(3000.0 / float(ours.sum()), 1), "xRT_reference_1core": round(3000.0 / float(refms.sum()), 1),
       "parity": {"vuv": M.vuv_agreement(r[0], o[0]), "f0_rel": M.f0_rel_error(r[0], o[0]),
                  "lsd_max_db": M.lsd_db(r[1], o[1])[1], "ap_abs": M.ap_abs_error(r[2], o[2])}}
print(json.dumps(res))

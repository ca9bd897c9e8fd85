"""BASELINE.json configs[3]: batched Synthesis-only resynthesis of 48 kHz f0 / sp / ap parameters.

The parameters are the synth tool's raw float32 files (W/test/synth.cpp:160-190: f0 [frames],
sp / ap [frames][fft_size/2+1]).  One batch of --utts utterances is analysed once on the device to
obtain realistic parameters, parked in pinned host memory as float32, and then resynthesised
--steps times; 10 h of audio is this batch repeated 36000 / audio_s times, so the rate is what matters.
  resident : Synthesis alone, parameters already in HBM (CUDA events on the library stream)
  e2e      : float32 parameters from pinned host memory -> device, Synthesis, 16-bit waveform back
Usage (GPU box): python scripts/config4_synthesis_only.py [--utts 300] [--steps 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hts_train_world_b200 as wb
from hts_train_world_b200 import signals

ap_ = argparse.ArgumentParser()
ap_.add_argument("--utts", type=int, default=300)
ap_.add_argument("--steps", type=int, default=3)
ap_.add_argument("--warmup", type=int, default=2)
args = ap_.parse_args()
FS = 48000
wb.init(0)
lengths = [int(round(signals.utterance_params(u)["T"] * FS)) for u in range(args.utts)]
pcm = torch.cat([signals.make_utterance(u, FS, device="cuda")[0] for u in range(args.utts)])
c = wb.Corpus(FS, lengths, 5.0)
c.set_pcm16_device(pcm)
c.analyze()
fft = c.fft_size
F, H = c.total_frames, fft // 2 + 1
f0_h = torch.from_numpy(c.f0().astype(np.float32)).pin_memory()
# sp / ap leave the device once and are kept as float32 (what the tool's files hold)
sp_h = torch.from_numpy(c.sp().astype(np.float32)).pin_memory()
ap_h = torch.from_numpy(c.ap().astype(np.float32)).pin_memory()
audio_s = sum(lengths) / FS
y_total = None
pcm_out = None
res, e2e = [], []
for it in range(args.warmup + args.steps):
    t0 = time.perf_counter()
    c.set_params_f32(fft, f0_h, sp_h, ap_h)
    c.synthesis()
    if pcm_out is None:
        pcm_out = torch.empty(int(wb.lib().wb200_batch_total_y(c._h)), dtype=torch.int16).pin_memory()
    c.y_pcm16(pcm_out)
    wb.sync()
    t1 = time.perf_counter()
    if it >= args.warmup:
        e2e.append(t1 - t0)
        res.append(wb.stage_times()["synthesis"] * 1e-3)
h2d = f0_h.numel() * 4 + sp_h.numel() * 4 + ap_h.numel() * 4
print(json.dumps({
    "config": "Synthesis-only, %d utterances (%.0f s of 48 kHz audio, %d frames) per batch, float32 f0/sp/ap in, 16-bit waveform out" % (args.utts, audio_s, F),
    "resident_xRT": round(audio_s / float(np.mean(res)), 1), "resident_ms_per_batch": round(1e3 * float(np.mean(res)), 2),
    "e2e_xRT": round(audio_s / float(np.mean(e2e)), 1), "e2e_ms_per_batch": round(1e3 * float(np.mean(e2e)), 2),
    "h2d_bytes_per_batch": h2d, "d2h_bytes_per_batch": pcm_out.numel() * 2,
    "h2d_GBps_needed_at_resident_rate": round(h2d / float(np.mean(res)) / 1e9, 1),
    "ten_hours_s": {"resident": round(36000.0 / (audio_s / float(np.mean(res))), 2), "e2e": round(36000.0 / (audio_s / float(np.mean(e2e))), 2)},
    "steps": args.steps, "warmup": args.warmup}))

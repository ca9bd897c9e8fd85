"""Generates the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/libworld_ref.so, compiled from /root/reference by oracle/Makefile) on
  * the reference's own two audio fixtures (externs/WORLD_v2/wav_test/arctic_a0001.wav,
    externs/WORLD_v2/test/vaiueo2d.wav), whose PCM is stored here so the tests can run where
    /root/reference does not exist,
  * two synthetic utterances of hts-train-world_b200/signals.py (16 kHz and 48 kHz).
Run in the build container:  python tests/golden/make_golden.py
Large arrays are stored as float32 (relative 6e-8: far below every tolerance) and the
spectrogram / aperiodicity only for every 8th frame, to keep the fixtures small.
"""
import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import hts_train_world_b200  # noqa: E402  (package shim)
from hts_train_world_b200 import signals  # noqa: E402
from oracle import ref  # noqa: E402

REFW = "/root/reference/externs/WORLD_v2"
ROW_STEP = 8


def read_wav(path):
    with wave.open(path, "rb") as w:
        assert w.getnchannels() == 1 and w.getsampwidth() == 2
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy(), w.getframerate()


def golden(name, pcm, fs, R):
    x = pcm.astype(np.float64) / 32768.0          # wavread, W/test/audioio.cpp:229-251
    a = R.analyze(x, fs)
    y = R.synthesis(a["f0"], a["sp"], a["ap"], a["fft_size"], 5.0, fs)
    _, f0_harvest = R.harvest(x, fs)
    rows = np.arange(0, len(a["f0"]), ROW_STEP)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), pcm=pcm, fs=fs, fft_size=a["fft_size"], t=a["t"],
        f0_raw=a["f0_raw"], f0=a["f0"], f0_harvest=f0_harvest, rows=rows,
        sp_rows=a["sp"][rows].astype(np.float32), ap_rows=a["ap"][rows].astype(np.float32),
        y=y.astype(np.float32), y_energy=float(np.sum(y * y)))
    print(name, fs, len(pcm), "frames", len(a["f0"]), "voiced", int((a["f0"] > 0).sum()))


if __name__ == "__main__":
    R = ref.load()
    pcm, fs = read_wav(os.path.join(REFW, "wav_test", "arctic_a0001.wav"))
    golden("arctic_a0001", pcm, fs, R)
    pcm, fs = read_wav(os.path.join(REFW, "test", "vaiueo2d.wav"))
    golden("vaiueo2d", pcm, fs, R)
    pcm, _ = signals.make_utterance(7, 48000, duration=1.5)
    golden("synthetic48k_u7", pcm.numpy(), 48000, R)
    pcm, _ = signals.make_utterance(11, 16000, duration=3.0)   # BASELINE.json configs[0]
    golden("synthetic16k_u11", pcm.numpy(), 16000, R)
    np.save(os.path.join(HERE, "randn_first_8192.npy"), ref.randn_stream(8192))

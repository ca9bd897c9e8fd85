import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hts_train_world_b200 as wb
from hts_train_world_b200 import signals
from oracle import ref, metrics as M
fs = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
R = ref.load(); wb.init(0)
pcm, _ = signals.make_utterance(0, fs, duration=2.0)
x = signals.pcm_to_double(pcm)
o = R.analyze(x, fs)
f0, fft = o["f0"], o["fft_size"]
yr = R.synthesis(f0, o["sp"], o["ap"], fft, 5.0, fs)
yn = wb.synthesis(f0, o["sp"], o["ap"], fft, 5.0, fs)
e = yr - yn
print("SNR", M.snr_db(yr, yn), "len", len(yr))
blk = 256
eb = np.add.reduceat(e * e, np.arange(0, len(e), blk))
sb = np.add.reduceat(yr * yr, np.arange(0, len(e), blk))
top = np.argsort(eb)[::-1][:12]
for b in sorted(top):
    s0 = b * blk
    fr = s0 / fs / 0.005
    print("block %5d sample %7d frame %.2f  err %.3e sig %.3e  f0 around %s" % (b, s0, fr, eb[b], sb[b], f0[max(0, int(fr) - 2):int(fr) + 3]))

"""Device time and HBM fraction of the cmp composer (delta windows + stream merge, wb_cmp.cu) on the
coded features of the bench corpus: the one HBM-bound kernel of the path.
Usage (GPU box): python scripts/cmp_throughput.py [--utts 1132]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hts_train_world_b200 as wb
from hts_train_world_b200 import signals

ap_ = argparse.ArgumentParser()
ap_.add_argument("--utts", type=int, default=1132)
args = ap_.parse_args()
FS = 48000
wb.init(0)
lengths = [int(round(signals.utterance_params(u)["T"] * FS)) for u in range(args.utts)]
c = wb.Corpus(FS, lengths, 5.0)
F = c.total_frames
rng = np.random.default_rng(0)
# the composer only reads the float32 statics: synthetic ones of the right shape are enough here
streams = [("mgc", rng.standard_normal((F, 50)).astype(np.float32)), ("lf0", rng.standard_normal((F, 1)).astype(np.float32)),
           ("bap", rng.standard_normal((F, 24)).astype(np.float32))]
arr = (wb.CmpStream * 3)()
for i, (_, data) in enumerate(streams):
    arr[i].source, arr[i].dim, arr[i].host_data, arr[i].n_win = 0, data.shape[1], data.ctypes.data, 3
    for w, coef in enumerate(wb.DEFAULT_WINDOWS):
        arr[i].win_size[w] = len(coef)
        for k, v in enumerate(coef):
            arr[i].win_coef[w][k] = v
wb.kernel_timing(True)
for _ in range(3):
    wb._check(wb.lib().wb200_batch_compose_cmp(c._h, arr, 3), "compose")
wb.kernel_times_reset()
n = 5
for _ in range(n):
    wb._check(wb.lib().wb200_batch_compose_cmp(c._h, arr, 3), "compose")
ms, launches = wb.kernel_time("cmp_compose_kernel")
ms /= launches
cmp_dim = int(wb.lib().wb200_batch_cmp_dim(c._h))
bytes_alg = F * (cmp_dim + 75) * 4
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm = float(peaks.get("hbm_gbs", 6650.0))
print(json.dumps({"kernel": "cmp_compose_kernel", "frames": F, "cmp_dim": cmp_dim, "ms_per_launch": round(ms, 4),
                  "algorithmic_bytes": bytes_alg, "achieved_GBps": round(bytes_alg / ms / 1e6, 1), "hbm_peak_GBps": hbm,
                  "frac": round(bytes_alg / ms / 1e6 / hbm, 3)}))

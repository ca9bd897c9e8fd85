"""Generates tests/golden/cmp_perl.npz by RUNNING THE REFERENCE'S OWN PERL SCRIPTS
(/root/reference/data/scripts/window.pl and addhtkheader.pl, with the window files of
/root/reference/data/win/) on seeded float32 streams.  Run in the build container:
    python tests/golden/make_golden_cmp.py
`merge` (SPTK) is not installed here; the side-by-side merge of data/Makefile.in:307-309 is the
concatenation [mgc | lf0 | bap | vib] per frame, done with numpy below and stated in the fixture.
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data"


def run_window(static, wins):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "in.f32")
        static.astype("<f4").tofile(p)
        out = subprocess.run(["perl", os.path.join(REF, "scripts", "window.pl"), str(static.shape[1]), p] + wins,
                             check=True, capture_output=True).stdout
    return np.frombuffer(out, "<f4").reshape(static.shape[0], -1).copy()


def run_header(samp, shift, byte, kind, payload):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "tmp.cmp")
        with open(p, "wb") as f:
            f.write(payload)
        return subprocess.run(["perl", os.path.join(REF, "scripts", "addhtkheader.pl"), str(samp), str(shift), str(byte),
                               str(kind), p], check=True, capture_output=True).stdout


if __name__ == "__main__":
    rng = np.random.default_rng(20261018)
    out = {}
    lengths = [1, 2, 3, 57, 200]                        # utterance lengths in frames (edge clamps matter)
    dims = {"mgc": 50, "lf0": 2, "bap": 25, "vib": 2}  # MGCORDER+1, LF0ORDER+1, BAPORDER+1, VIBORDER+1
    out["lengths"] = np.array(lengths)
    for name, dim in dims.items():
        wins = [os.path.join(REF, "win", "%s.win%d" % (name, i)) for i in (1, 2, 3)]
        for u, T in enumerate(lengths):
            s = rng.standard_normal((T, dim)).astype(np.float32) * 3.0
            if name in ("lf0", "vib") and T > 10:       # unvoiced stretches carrying the ignore value
                s[5:9, 0] = -1.0e10
                s[T - 2:, 0] = -1.0e10
                s[20:21, 1] = -1.0e10
            out["%s_static_%d" % (name, u)] = s
            out["%s_windowed_%d" % (name, u)] = run_window(s, wins)
    # a window with leading / trailing zero taps (boundary check skips them, window.pl:70-81)
    with tempfile.TemporaryDirectory() as d:
        wf = os.path.join(d, "odd.win")
        with open(wf, "w") as f:
            f.write("5 0.0 -0.25 0.5 0.75 0.0\n")
        s = out["lf0_static_3"]
        out["odd_window"] = np.array([0.0, -0.25, 0.5, 0.75, 0.0])
        out["odd_windowed"] = run_window(s, [wf])
    # HTK header in front of one composed utterance
    u = 3
    cmp_u = np.concatenate([out["%s_windowed_%d" % (n, u)] for n in ("mgc", "lf0", "bap", "vib")], axis=1)
    byte = 4 * cmp_u.shape[1]
    blob = run_header(48000, 240, byte, 9, cmp_u.astype("<f4").tobytes())
    out["htk_header"] = np.frombuffer(blob[:12], np.uint8).copy()
    out["htk_args"] = np.array([lengths[u], 48000, 240, byte, 9])
    assert blob[12:] == cmp_u.astype("<f4").tobytes()
    np.savez_compressed(os.path.join(HERE, "cmp_perl.npz"), **out)
    print("wrote cmp_perl.npz:", {k: v.shape for k, v in out.items() if k.startswith("mgc_windowed")})

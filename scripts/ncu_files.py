"""Executed warp instructions / stall samples of one kernel summed per source file and per line range (run here).
  python scripts/ncu_files.py gpurun_out/prof.ncu-rep d4c_gd_kernel
"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr = None, None
inst = collections.Counter(); smp = collections.Counter(); lines = {}
nfile = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and fpath:
        d = dict(zip(hdr, r))
        try:
            n = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
        except (ValueError, KeyError):
            continue
        inst[fpath] += n; smp[fpath] += s
        lines[(fpath, int(r[0]))] = lines.get((fpath, int(r[0])), 0) + n
ti = sum(inst.values()) or 1; ts = sum(smp.values()) or 1
for f, n in inst.most_common():
    print("%-28s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * n / ti, 100.0 * smp[f] / ts))
if len(sys.argv) > 3:
    f = sys.argv[3]
    step = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    buckets = collections.Counter()
    for (ff, ln), n in lines.items():
        if ff == f: buckets[ln // step * step] += n
    for b in sorted(buckets):
        if buckets[b] * 1000 > ti: print("  %s:%d-%d  %5.1f%%" % (f, b, b + step - 1, 100.0 * buckets[b] / ti))

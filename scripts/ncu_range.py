"""Per-line executed warp instructions and stall samples for a source line range of one kernel.
  python scripts/ncu_range.py rep kernel file first last"""
import csv, io, subprocess, sys
rep, kern, fname, a, b = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr = None, None
tot = 0; acc = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and fpath:
        d = dict(zip(hdr, r))
        try: n = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
        except (ValueError, KeyError): continue
        tot += n
        if fpath == fname and a <= int(r[0]) <= b:
            e = acc.setdefault(int(r[0]), [0, 0, r[1]]); e[0] += n; e[1] += s
for ln in sorted(acc):
    n, s, src = acc[ln]
    print("%5d %6.2f%% %7d  %s" % (ln, 100.0 * n / tot, s, src[:150]))

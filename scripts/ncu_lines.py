"""Per-source-line hot spots of one kernel from an ncu report (run here).
  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep d4c_main_kernel [top]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sort_key = {"samples": 3, "inst": 4, "shared": 5}[sys.argv[4] if len(sys.argv) > 4 else "samples"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr, items, seen_fn = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            items.append((fpath, int(r[0]), r[1].strip(), int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0),
                          int(d.get("L1 Wavefronts Shared") or 0), int(d.get("L1 Wavefronts Shared Ideal") or 0)))
        except ValueError:
            pass
ts = sum(i[3] for i in items) or 1
ti = sum(i[4] for i in items) or 1
print("total samples %d, warp instructions %d, shared wavefronts %d (ideal %d)" % (ts, ti, sum(i[5] for i in items), sum(i[6] for i in items)))
print("%-22s %5s %6s %6s %9s  %s" % ("file", "line", "smpl%", "inst%", "shm x/id", "source"))
for it in sorted(items, key=lambda x: -x[sort_key])[:top]:
    print("%-22s %5d %5.1f%% %5.1f%% %4.1f %5.1f%% %s" % (it[0][:22], it[1], 100.0 * it[3] / ts, 100.0 * it[4] / ti,
                                                   (it[5] / it[6]) if it[6] else 0.0, 100.0 * it[5] / max(1, sum(i[5] for i in items)), it[2][:100]))

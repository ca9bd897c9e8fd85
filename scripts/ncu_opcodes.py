"""Executed warp instructions of one kernel grouped by SASS opcode (run here).
  python scripts/ncu_opcodes.py gpurun_out/prof.ncu-rep d4c_gd_kernel [top]
"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
ops = collections.Counter()
smp = collections.Counter()
launches = 0
for r in rows:
    if not r:
        continue
    if r[0] in ("Address", "#"):
        hdr = r
        launches += 1
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        src = d.get("Source", "")
        try:
            n = int(d.get("Instructions Executed") or 0)
            s = int(d.get("# Samples") or 0)
        except ValueError:
            continue
        toks = src.replace("{", " ").split()
        toks = [t for t in toks if not t.startswith("@")]
        if not toks:
            continue
        op = toks[0].split(".")[0]
        ops[op] += n
        smp[op] += s
tot = sum(ops.values()) or 1
ts = sum(smp.values()) or 1
print("kernel %s: %d launches in report, %d warp instructions, %d samples" % (kern, launches, tot, ts))
for op, n in ops.most_common(top):
    print("  %-10s %6.2f%% inst  %6.2f%% samples" % (op, 100.0 * n / tot, 100.0 * smp[op] / ts))

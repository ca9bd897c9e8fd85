"""profiles/traffic.json from an ncu --set full report (run here):
  python scripts/ncu_traffic.py gpurun_out/r2h_prof_d4c_1132.ncu-rep [more.ncu-rep ...] --utts 1132 --out profiles/traffic.json
For every kernel in the reports: dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured
launches).  bench.py copies the entry of the dominant kernel into roofline.traffic; the file records which
report, which utterance count and which commit the numbers come from."""
import argparse
import csv
import io
import json
import re
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("reports", nargs="+")
ap.add_argument("--utts", type=int, required=True, help="utterances of the captured run (traffic scales with it)")
ap.add_argument("--out", default="profiles/traffic.json")
args = ap.parse_args()
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
acc = {}
for rep in args.reports:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"<.*", "", r[ix["Kernel Name"]].split("(")[0].replace("void ", "")).split("::")[-1].strip()
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)
        dur = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
        e = acc.setdefault(name, dict(bytes=0.0, n=0, report=rep.split("/")[-1], ms=0.0))
        e["bytes"] += tot
        e["ms"] += dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0)
        e["n"] += 1
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
res = {"_provenance": {"utterances": args.utts, "commit": head, "metric": "dram__bytes_read.sum + dram__bytes_write.sum per launch",
                       "note": "ncu replays each kernel with cold caches; durations under ncu are not bench values"}}
for k, e in sorted(acc.items()):
    res[k] = e["bytes"] / e["n"]
    res["_provenance"][k] = {"launches": e["n"], "report": e["report"], "ms_per_launch_under_ncu": e["ms"] / e["n"]}
json.dump(res, open(args.out, "w"), indent=1)
print(json.dumps(res, indent=1)[:1200])

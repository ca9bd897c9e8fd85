"""Writes profiles/sass/: gzip'd SASS of the heavy kernels of libworld_b200.so and their
instruction-mix histograms (cuobjdump -sass; run here, no GPU needed)."""
import collections
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hts-train-world_b200", "libworld_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
WANT = {
    "d4c_main_kernel<12,256,4>": r"d4c_main_kernelILi12ELi256ELi4E",
    "d4c_lovetrain_kernel<12>": r"d4c_lovetrain_kernelILi12E",
    "cheaptrick_kernel<11,128>": r"cheaptrick_kernelILi11ELi128E",
    "synth_item_kernel<11,float2,256>": r"synth_item_kernelILi11E6float2Li256E",
    "stonemask_kernel": r"stonemask_kernelE",
    "ols_filter_kernel<13>": r"wb_dio.*ols_filter_kernelILi13E|ols_filter_kernelILi13E",
    "harvest_refine_kernel": r"harvest_refine_kernelE",
    "codec_encode_kernel<10,fp32 log>": r"codec_encode_kernelILi10ELb1E",
}
syms = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
names = sorted(set(re.findall(r"\.text\.(_Z\w+)", syms)))
os.makedirs(OUT, exist_ok=True)
mix_lines = []
for label, pat in WANT.items():
    hit = [n for n in names if re.search(pat, n)]
    if not hit:
        print("no symbol for", label, file=sys.stderr)
        continue
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", hit[0], LIB], capture_output=True, text=True).stdout
    fn = re.sub(r"[^A-Za-z0-9_]+", "_", label).strip("_") + ".sass.gz"
    with gzip.open(os.path.join(OUT, fn), "wt") as f:
        f.write(sass)
    ops = collections.Counter(re.sub(r"\..*", "", m) for m in re.findall(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", sass, re.M))
    total = sum(ops.values())
    top = ", ".join("%s %d" % kv for kv in ops.most_common(14))
    mix_lines.append("%-36s %6d instructions: %s" % (label, total, top))
open(os.path.join(OUT, "instruction_mix.txt"), "w").write(
    "Static SASS instruction mix (cuobjdump -sass, sm_100a).  DADD/DMUL/DFMA = FP64 pipe, FADD/FMUL/FFMA = FP32,\n"
    "LDS/STS = shared-memory FFT traffic, ATOMS = selection histogram, MUFU = transcendental seeds, SHFL = warp scans.\n"
    "No HMMA/UTC*MMA: no stage is a dense contraction (BASELINE.json north_star).\n\n" + "\n".join(mix_lines) + "\n")
print("\n".join(mix_lines))

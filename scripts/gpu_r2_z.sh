#!/bin/bash
# Developer tool (GPU box), round 2 call Z: the coded features' copy queued behind Synthesis instead of in front of it
mkdir -p gpurun_out
for sk in "" late; do
WB_E2E_SKIP=$sk timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2z.json 2> gpurun_out/r2z.err; 
python - "$sk" <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2z.json'))
print("skip [%s]: resident %.2f ms  e2e %.2f ms" % (sys.argv[1], d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done

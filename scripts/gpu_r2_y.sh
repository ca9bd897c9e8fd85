#!/bin/bash
# Developer tool (GPU box), round 2 call Y: what each copy of the end-to-end leg costs (one left out at a time)
mkdir -p gpurun_out
for sk in "" up coded wave; do
WB_E2E_SKIP=$sk timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2y.json 2> gpurun_out/r2y.err; 
python - "$sk" <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2y.json'))
print("skip [%s]: resident %.2f ms  e2e %.2f ms" % (sys.argv[1], d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done

#!/bin/bash
# Developer tool (GPU box): what the end-to-end leg loses against the resident leg, by number of pipelined sub-batches
mkdir -p gpurun_out
for p in 1 2 3; do
  WB_E2E_PARTS=$p timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2j_parts$p.json 2> gpurun_out/r2j_parts$p.err
  python - $p <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2j_parts%s.json' % sys.argv[1]))
print("parts %s: resident %.1f ms  e2e %.1f ms (device %.1f, wall %.1f)  value %.0f e2e %.0f" % (sys.argv[1], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['device_ms_per_step'], d['e2e']['wall_ms_per_step'], d['value'], d['e2e']['value']))
PY
done

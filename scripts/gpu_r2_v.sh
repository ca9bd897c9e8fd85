#!/bin/bash
# Developer tool (GPU box), round 2 call V: where the end-to-end leg's extra milliseconds go (per-kernel timers during the e2e leg)
mkdir -p gpurun_out
for p in 2 1; do
WB_E2E_PARTS=$p WB_E2E_KTIME=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2v_p$p.json 2> gpurun_out/r2v_p$p.err; echo "bench exit $?"
grep "\[e2e\]" gpurun_out/r2v_p$p.err
python - $p <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2v_p%s.json' % sys.argv[1]))
print("parts %s: resident %.2f ms (kernels %.2f)  e2e %.2f ms" % (sys.argv[1], d['ms_per_step'], sum(v['ms_per_launch'] * v['launches_per_step'] for v in d['kernels'].values()), d['e2e']['ms_per_step']), d['stage_ms'])
PY
done

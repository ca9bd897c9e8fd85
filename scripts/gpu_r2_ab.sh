#!/bin/bash
# Developer tool (GPU box), round 2 call AB: randn table sized by a host-side bound (no read-backs in CheapTrick / D4C), deferral test: parity, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2ab_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2ab.json 2> gpurun_out/r2ab.err; tail -2 gpurun_out/r2ab.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ab.json'))
print("resident %.2f ms  e2e %.2f ms  (%.0f / %.0f xRT) parity %s launches %s" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['within_tolerance'], d['gpu_launches']))
PY

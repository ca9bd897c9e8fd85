#!/bin/bash
# Developer tool (GPU box), round 2 call X: small host -> device tables through mapped memory (write_dev): parity, end-to-end leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2x_pytest.log
WB_E2E_KTIME=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --verify 2 > gpurun_out/r2x.json 2> gpurun_out/r2x.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2x.json'))
print("resident %.2f ms (kernels %.2f)  e2e %.2f ms  value %.0f e2e %.0f parity %s" % (d['ms_per_step'], sum(v['ms_per_launch'] * v['launches_per_step'] for v in d['kernels'].values()), d['e2e']['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['within_tolerance']))
PY
timeout 900 python bench.py --f0 harvest --steps 3 --warmup 3 --no-cpu-baseline --no-configs --verify 2 > gpurun_out/r2x_harvest.json 2> gpurun_out/r2x_harvest.err; echo "harvest bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2x_harvest.json'))
print("harvest: resident %.2f ms e2e %.2f ms value %.0f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value']), d['parity'])
PY

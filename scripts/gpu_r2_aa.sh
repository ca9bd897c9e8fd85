#!/bin/bash
# Developer tool (GPU box), round 2 call AA: deferred bulk copies (wb200_set_copy_deferral): parity, end-to-end leg with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2aa_pytest.log
for d in 1 0; do
WB_E2E_DEFER=$d timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2aa.json 2> gpurun_out/r2aa.err; tail -2 gpurun_out/r2aa.err
python - "$d" <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2aa.json'))
print("defer %s: resident %.2f ms  e2e %.2f ms  (%.0f / %.0f xRT)" % (sys.argv[1], d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['e2e']['value']))
PY
done
rm -rf /tmp/cfg5 && mkdir -p /tmp/cfg5
timeout 900 python hts-train-world_b200/driver.py --synthetic-hours 4 --out-dir /tmp/cfg5 --fs 48000 2>&1 | tail -2

#!/bin/bash
# Developer tool (GPU box), round 2 call L: packed-bf16 selection, constant-offset smoothing, FP32 row: parity, A/B, step trace of the Harvest configuration
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2l_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "WB200_LIB=libworld_b200_base.so" "WB200_LIB=libworld_b200_sel32.so" "" "WB200_LIB=libworld_b200_sel32.so" ""
WB_STEP_TRACE=1 timeout 900 python bench.py --f0 harvest --steps 2 --warmup 1 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2l_harvest.json 2> gpurun_out/r2l_harvest.err; echo "harvest bench exit $?"
grep "\[step\]" gpurun_out/r2l_harvest.err | tail -20
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2l_harvest.json'))
print("harvest full: value %.0f ms %.2f" % (d['value'], d['ms_per_step']), d.get('stage_ms'))
PY

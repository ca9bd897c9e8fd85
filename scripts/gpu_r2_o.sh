#!/bin/bash
# Developer tool (GPU box), round 2 call O: twiddles formed before the pass barrier, Goertzel recurrences in the Harvest refinement
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2o_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "WB200_LIB=libworld_b200_base.so" "" "WB200_LIB=libworld_b200_sel32.so" ""
timeout 900 python bench.py --f0 harvest --steps 3 --warmup 3 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2o_harvest.json 2> gpurun_out/r2o_harvest.err; echo "harvest bench exit $?"
python - <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2o_harvest.json'))
print("harvest full: value %.0f ms %.2f | " % (d['value'], d['ms_per_step']) + " ".join("%s %.1f" % (k.replace('_kernel', ''), v['ms_per_launch'] * v['launches_per_step']) for k, v in d['kernels'].items()), d.get('stage_ms'))
PY

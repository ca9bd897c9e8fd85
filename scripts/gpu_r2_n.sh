#!/bin/bash
# Developer tool (GPU box), round 2 call N: warp-level DC correction, cumsum without the protective barrier, grouped
# single-precision sums in the selection, CheapTrick's constant-offset smoothing; Harvest with grow-only sub-batch scratch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2n_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "WB200_LIB=libworld_b200_base.so" "" "WB200_LIB=libworld_b200_sel32.so" ""
for w in 1 3; do
timeout 900 python bench.py --f0 harvest --steps 3 --warmup $w --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2n_harvest_w$w.json 2> gpurun_out/r2n_harvest_w$w.err; echo "harvest bench exit $?"
python - $w <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2n_harvest_w%s.json' % sys.argv[1]))
print("harvest full (warmup %s): value %.0f ms %.2f | " % (sys.argv[1], d['value'], d['ms_per_step']) + " ".join("%s %.1f" % (k.replace('_kernel', ''), v['ms_per_launch'] * v['launches_per_step']) for k, v in d['kernels'].items()), d.get('stage_ms'))
PY
done

#!/bin/bash
# Developer tool (GPU box), round 2 call K: parity suite on the FMA-form butterflies / warp-local pass barrier,
# A/B against the previous library (libworld_b200_base.so), full-size line, phase trace of Harvest at 1 132 utterances.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2k_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "WB200_LIB=libworld_b200_base.so" "" "WB200_LIB=libworld_b200_base.so" ""
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2k_bench.json'))
print("full: value %.0f ms %.2f e2e %.2f | " % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']) + " ".join("%s %.1f" % (k.replace('_kernel', ''), v['ms_per_launch'] * v['launches_per_step']) for k, v in d['kernels'].items()))
print("parity", d.get('parity', {}).get('within_tolerance'), {k: d['parity'][k] for k in ('f0_rel_error', 'lsd_db_max', 'ap_abs_error', 'snr_db')})
PY
WB_HARVEST_TRACE=1 timeout 900 python bench.py --f0 harvest --steps 2 --warmup 1 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2k_harvest.json 2> gpurun_out/r2k_harvest.err; echo "harvest bench exit $?"
grep "\[harvest\]" gpurun_out/r2k_harvest.err | tail -40
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2k_harvest.json'))
print("harvest full: value %.0f ms %.2f | " % (d['value'], d['ms_per_step']) + " ".join("%s %.1f" % (k.replace('_kernel', ''), v['ms_per_launch'] * v['launches_per_step']) for k, v in d['kernels'].items()), d.get('stage_ms'))
PY

#!/bin/bash
# Developer tool (GPU box), round 2 call M: full-size line, Harvest configuration with bounded segment scratch (verified
# against the reference's Harvest), --set full capture with source of the new d4c_main / cheaptrick / synth_item at 64 utterances
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2m_bench.json'))
print("full: value %.0f ms %.2f e2e %.2f | " % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']) + " ".join("%s %.1f" % (k.replace('_kernel', ''), v['ms_per_launch'] * v['launches_per_step']) for k, v in d['kernels'].items()))
print("parity", d.get('parity', {}).get('within_tolerance'), {k: d['parity'][k] for k in ('f0_rel_error', 'lsd_db_max', 'ap_abs_error', 'snr_db')})
print("configs", json.dumps(d.get('configs'))[:1500])
PY
timeout 900 python bench.py --f0 harvest --steps 3 --warmup 3 --no-cpu-baseline --no-configs --verify 2 > gpurun_out/r2m_harvest.json 2> gpurun_out/r2m_harvest.err; echo "harvest bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2m_harvest.json'))
print("harvest full: value %.0f ms %.2f | " % (d['value'], d['ms_per_step']) + " ".join("%s %.1f" % (k.replace('_kernel', ''), v['ms_per_launch'] * v['launches_per_step']) for k, v in d['kernels'].items()), d.get('stage_ms'))
print("parity", d.get('parity'))
PY
SMALL="python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --verify 0"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:d4c_main|synth_item|cheaptrick_kernel|lovetrain|ols_filter" -c 6 -f -o gpurun_out/r2m_prof $SMALL > gpurun_out/r2m_ncu.log 2>&1; echo "ncu exit $?"

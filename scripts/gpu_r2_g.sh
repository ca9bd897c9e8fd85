#!/bin/bash
# Developer tool (GPU box), round 2 call G: the whole parity suite (thread-safety lock, private memory pool,
# pipelined driver with resume, GV statistics, shared window walks in the Harvest refinement), A/B lines on
# both F0 paths, one full default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2g_pytest.log
UTTS=300 bash scripts/gpu_ab.sh ""
BENCH_EXTRA="--f0 harvest" UTTS=300 bash scripts/gpu_ab.sh ""
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2g_bench.json'))
print("value %.0f e2e %.0f ms/step %.1f e2e ms %.1f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
print("stage_ms", {k: round(v, 1) for k, v in d['stage_ms'].items()})
for k, v in d['kernels'].items():
    print("  %-24s %8.2f ms x%.0f  %s" % (k, v['ms_per_launch'], v['launches_per_step'], ("frac %.3f (%s)" % (v['frac'], v['bound'])) if 'frac' in v else ""))
print("roofline", {k: d['roofline'][k] for k in ('kernel', 'bound', 'achieved', 'peak', 'frac')})
print("roofline_step", {k: d['roofline_step'][k] for k in ('bound', 'achieved', 'peak', 'frac')})
print("parity", {k: d['parity'][k] for k in ('vuv_agreement', 'f0_rel_error', 'lsd_db_max', 'ap_abs_error', 'snr_db', 'within_tolerance')})
print("configs", json.dumps(d['configs'])[:1800])
print("cpu", d.get('cpu_baseline', {}).get('value'), d['clocks'])
PY

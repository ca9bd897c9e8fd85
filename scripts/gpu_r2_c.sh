#!/bin/bash
# Developer tool (GPU box), round 2 call C: parity suite with the fused Dio / Harvest filters and the
# thread-per-candidate Harvest refinement as defaults; A/B bench lines (Dio path and Harvest path);
# one full default bench line (verify + configs legs).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2c_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "" "WB_DIO_FUSED=0"
BENCH_EXTRA="--f0 harvest" UTTS=300 bash scripts/gpu_ab.sh "" "WB_HARVEST_FUSED=0" "WB_HARVEST_REFINE_THREAD=0"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_bench.json'))
print("value %.0f e2e %.0f ms/step %.1f" % (d['value'], d['e2e']['value'], d['ms_per_step']))
print("stage_ms", {k: round(v, 1) for k, v in d['stage_ms'].items()})
for k, v in d['kernels'].items():
    print("  %-24s %8.2f ms x%.0f  %s" % (k, v['ms_per_launch'], v['launches_per_step'], ("frac %.3f (%s)" % (v['frac'], v['bound'])) if 'frac' in v else ""))
print("roofline", {k: d['roofline'][k] for k in ('kernel', 'bound', 'achieved', 'peak', 'frac')})
print("roofline_step", {k: d['roofline_step'][k] for k in ('bound', 'achieved', 'peak', 'frac')})
print("parity", d['parity'])
print("configs", json.dumps(d['configs'])[:1500])
print("cpu", d.get('cpu_baseline', {}).get('value'), d['clocks'])
PY

#!/bin/bash
# Developer tool (GPU box): parity tests + one bench line without the CPU baseline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value %.0f e2e %.0f ms/step %.1f launches %d"%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches']))
print("stage_ms",{k:round(v,1) for k,v in d['stage_ms'].items()})
for k,v in d['kernels'].items(): print("  %-24s %8.2f ms x%.0f  %s"%(k,v['ms_per_launch'],v['launches_per_step'],("fp64 frac %.3f"%v['frac_fp64']) if 'frac_fp64' in v else ""))
print(d['roofline']['kernel'],d['roofline']['frac'],d['clocks'])
PY

#!/bin/bash
# Developer tool (GPU box): what the driver runs at round end -- the GPU parity suite, smoke(), a short default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/final_bench.json'))
print("value %.0f ms %.2f e2e %.0f ms %.2f frac %.4f parity %s cpu %.1f launches %d clocks %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity']['within_tolerance'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']))
PY

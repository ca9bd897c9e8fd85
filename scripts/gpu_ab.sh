#!/bin/bash
# Developer tool (GPU box): A/B a list of environment settings on a reduced bench run.
# Usage: gpurun -- 'bash scripts/gpu_ab.sh "" "WB_D4C_T512=1"'
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --no-cpu-baseline --no-configs --verify 0 ${BENCH_EXTRA} --utts ${UTTS:-300} --steps 3 --warmup 2 > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err || tail -5 gpurun_out/ab_$i.err
  python - "$envs" gpurun_out/ab_$i.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2]))
print("[%s] value %.0f ms/step %.1f | "%(sys.argv[1],d['value'],d['ms_per_step'])+" ".join("%s %.1f"%(k.replace('_kernel',''),v['ms_per_launch']*v['launches_per_step']) for k,v in d['kernels'].items()))
PY
done

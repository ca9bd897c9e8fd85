#!/bin/bash
# Developer tool (8-GPU box): the weak-scaling bench line at N = 8 (deferred copies on / off through WB_E2E_DEFER)
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N exit $?"; tail -2 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_n$N.json'))
print("N=%d value %.0f e2e %.0f ms/step %.1f e2e ms %.1f deferred %s" % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('deferred_copies')))
PY

#!/bin/bash
# Developer tool (GPU box): the parity suite and a reduced bench line inside a ~40 s budget.
mkdir -p gpurun_out
timeout 27 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
timeout 14 python bench.py --no-cpu-baseline --utts 300 --steps 2 --warmup 1 > gpurun_out/bench_e2e_fix.json 2> gpurun_out/bench_e2e_fix.err
echo "bench exit $?"
cut -c1-900 gpurun_out/bench_e2e_fix.json

#!/bin/bash
# Developer tool (GPU box): parity tests, the default bench line, and the ncu launch list of a
# small bench run.  Usage: gpurun --timeout 1800 -- 'bash scripts/gpu_check.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
SMALL="python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $SMALL > gpurun_out/plain_small.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_small.log 2>&1
echo "ncu exit $?"

#!/bin/bash
# Developer tool (GPU box): ncu launch list of our kernels and one --set full capture of each
# heavy kernel on a small bench run.  Usage: gpurun --timeout 1800 -- 'bash scripts/gpu_ncu.sh'
set -x
mkdir -p gpurun_out
SMALL="python bench.py --utts ${UTTS:-64} --steps 1 --warmup 1 --no-cpu-baseline"
KRE='regex:dio_|d4c_|cheaptrick|synth_|stonemask|seg_scan|pcm16|lf0_stats|default_frames|randn_table|harvest|ols_filter|zc_|codec'
HEAVY="regex:${HEAVY:-d4c_main|synth_item|cheaptrick_kernel|stonemask_kernel|lovetrain|ols_filter}"
timeout 600 $SMALL > gpurun_out/plain_small.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_small.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k "$HEAVY" -c ${NCAP:-6} -f -o gpurun_out/prof $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
tail -3 gpurun_out/ncu_full.log

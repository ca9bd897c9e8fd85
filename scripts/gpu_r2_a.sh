#!/bin/bash
# Developer tool (GPU box), round 2 first call: the whole parity suite (new: other sampling rates, hard
# inputs), then with the split D4C kernels + FP32 LoveTrain switched on; A/B bench lines; one default
# bench line (verify + configs legs); ncu --set full of the heavy kernels with the split kernels on.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_r2_a.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2a_pytest.log
WB_D4C_SPLIT=1 WB_D4C_LT32=1 timeout 600 python -m pytest tests -m gpu -q -k "d4c or stages or end_to_end or batch or silence or short or extremes or full_size or hard or sampling" > gpurun_out/r2a_pytest_split.log 2>&1
echo "split pytest exit $?"; tail -8 gpurun_out/r2a_pytest_split.log
UTTS=300 bash scripts/gpu_ab.sh "" "WB_D4C_SPLIT=1" "WB_D4C_LT32=1" "WB_D4C_SPLIT=1 WB_D4C_LT32=1"
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2a_bench.err
cut -c1-1500 gpurun_out/r2a_bench.json
SMALL="python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --verify 0"
WB_D4C_SPLIT=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:d4c_gd|d4c_tail|d4c_main|synth_item|cheaptrick_kernel|stonemask_kernel|lovetrain|ols_filter|zc_kernel|codec_encode" -c 12 -f -o gpurun_out/r2a_prof $SMALL > gpurun_out/r2a_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r2a_ncu.log

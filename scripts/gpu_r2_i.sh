#!/bin/bash
# Developer tool (GPU box), round 2 call I: the config-5 driver on one GPU (2 hours of synthetic audio, with and
# without files), A/B of two experimental builds, the --set full capture of the dominant kernel at 1 132 utterances.
mkdir -p gpurun_out
N=1 HOURS=2 bash scripts/gpu_r2_multi.sh 2>&1 | grep -v "^$" | tail -14
UTTS=300 bash scripts/gpu_ab.sh "" "WB200_LIB=libworld_b200_x1.so"
BENCH_EXTRA="--f0 harvest" UTTS=300 bash scripts/gpu_ab.sh "" "WB200_LIB=libworld_b200_x1.so"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:d4c_main_kernel" -c 1 -f -o gpurun_out/r2i_prof_d4c_1132 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2i_ncu_d4c.log 2>&1; echo "ncu d4c_main (1132 utts) exit $?"

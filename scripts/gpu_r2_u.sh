#!/bin/bash
# Developer tool (GPU box), round 2 call U: experimental builds: synthesis items with 128-thread CTAs / radix-8 passes, CheapTrick's FP32 transforms with radix-8 passes
mkdir -p gpurun_out
UTTS=300 bash scripts/gpu_ab.sh "" "WB200_LIB=libworld_b200_t128.so" "WB200_LIB=libworld_b200_k3.so" "WB200_LIB=libworld_b200_ct3.so" ""

#!/bin/bash
# Developer tool (GPU box), round 2 call AE: L2 prefetch of a frame's dither words at the start of d4c_main
mkdir -p gpurun_out
UTTS=300 bash scripts/gpu_ab.sh "WB200_LIB=libworld_b200_nopf.so" "" "WB200_LIB=libworld_b200_nopf.so" ""

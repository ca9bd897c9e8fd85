#!/bin/bash
# Developer tool (GPU box), round 2 call D: parity suite (pruned D4C sweeps, integer overlap-add, warp-cooperative
# Harvest contour logic, Harvest segments of 256); A/B bench lines on both F0 paths.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2d_pytest.log
UTTS=300 bash scripts/gpu_ab.sh ""
BENCH_EXTRA="--f0 harvest" UTTS=300 bash scripts/gpu_ab.sh "" "WB_HARVEST_FUSED=0"

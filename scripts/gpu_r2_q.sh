#!/bin/bash
# Developer tool (GPU box), round 2 call Q: LoveTrain with the TMA-staged window: parity, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2q_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "WB200_LIB=libworld_b200_base.so" "" "WB200_LIB=libworld_b200_base.so" ""

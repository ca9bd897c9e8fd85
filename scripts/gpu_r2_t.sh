#!/bin/bash
# Developer tool (GPU box), round 2 call T: synthesis time base with four samples per thread: parity, A/B against one sample per thread
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2t_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "WB_SYNTH_PHASE4=0" "" "WB_SYNTH_PHASE4=0" ""

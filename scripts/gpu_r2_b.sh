#!/bin/bash
# Developer tool (GPU box), round 2 call B: parity suite with the direct-DFT StoneMask and FP32 LoveTrain as
# defaults; the same suite on the experimental builds (twiddle squaring, dither preload); A/B bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2b_pytest.log
WB200_LIB=libworld_b200_both.so timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_both.log 2>&1; echo "pytest(both) exit $?"; tail -6 gpurun_out/r2b_pytest_both.log
UTTS=300 bash scripts/gpu_ab.sh "" "WB_STONEMASK_DFT=0" "WB200_LIB=libworld_b200_twsq.so" "WB200_LIB=libworld_b200_rnpre.so" "WB200_LIB=libworld_b200_both.so" ""

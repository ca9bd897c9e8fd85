#!/bin/bash
# Developer tool (GPU box), round 2 call E: which of the two Harvest changes broke three Harvest tests?
mkdir -p gpurun_out
WB_HARVEST_FUSED=0 timeout 600 python -m pytest tests -m gpu -q -k "harvest or long_utterance" > gpurun_out/r2e_pytest_nofuse.log 2>&1; echo "harvest tests, two-pass filter: exit $?"; tail -4 gpurun_out/r2e_pytest_nofuse.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2e_pytest.log

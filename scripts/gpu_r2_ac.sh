#!/bin/bash
# Developer tool (GPU box), round 2 call AC: window phasors computed once per frame in d4c_main: parity, reduced line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2ac_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "" ""

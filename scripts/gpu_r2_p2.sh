#!/bin/bash
# Developer tool (GPU box), round 2 call P2: --set full of the dominant kernel at the full 1 132 utterances (its own call: the reports of P and P2 together exceed what one call may bring back)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:d4c_main_kernel" -c 1 -f -o gpurun_out/r2p_prof_d4c_1132 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2p_ncu_d4c.log 2>&1; echo "ncu d4c_main (1132 utts) exit $?"

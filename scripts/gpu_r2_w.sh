#!/bin/bash
# Developer tool (GPU box), round 2 call W: wall time of every call of one end-to-end step
mkdir -p gpurun_out
WB_E2E_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --verify 0 > gpurun_out/r2w.json 2> gpurun_out/r2w.err; echo "bench exit $?"
grep "\[e2e step\]" gpurun_out/r2w.err | awk '{s+=$(NF-1); print} END{print "sum", s}'

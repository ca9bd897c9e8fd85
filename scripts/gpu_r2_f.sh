#!/bin/bash
mkdir -p gpurun_out
WB_HARVEST_FIX_CHECK=3 timeout 600 python -m pytest tests -m gpu -q -s -k "harvest_golden or long_utterance" 2>&1 | grep -E "fix check\] [0-9]+ frames|passed|failed" | head -10
timeout 600 python -m pytest tests -m gpu -q -k "harvest or long_utterance or sampling" 2>&1 | tail -3

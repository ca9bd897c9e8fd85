#!/bin/bash
# Developer tool (GPU box), round 2 call P: parity suite, default bench line, then the profile evidence of the
# round: ncu launch list of a small bench run, --set full of every heavy kernel (64 utterances), --set full of
# the dominant kernel at the full 1 132 utterances.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2p_pytest.log
UTTS=300 bash scripts/gpu_ab.sh "" ""
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2p_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2p_bench.json'))
print("value %.0f e2e %.0f ms/step %.1f e2e ms %.1f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
for k, v in d['kernels'].items():
    print("  %-24s %8.2f ms x%.0f  %s" % (k, v['ms_per_launch'], v['launches_per_step'], ("frac %.3f (%s)" % (v['frac'], v['bound'])) if 'frac' in v else ""))
print("configs", json.dumps(d['configs'])[:1500])
PY
SMALL="python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --verify 0"
KRE='regex:dio_|d4c_|cheaptrick|synth_|stonemask|seg_scan|pcm16|lf0_|stats|default_frames|randn_table|ols_filter|zc_|codec|gv_|feature_|widen|read_back'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/r2p_launches.csv $SMALL > gpurun_out/r2p_ncu_small.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:d4c_main|synth_item|cheaptrick_kernel|stonemask_dft|lovetrain|ols_filter_zc|codec_encode|synth_phase" -c 10 -f -o gpurun_out/r2p_prof $SMALL > gpurun_out/r2p_ncu_full.log 2>&1; echo "ncu full (64 utts) exit $?"

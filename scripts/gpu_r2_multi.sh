#!/bin/bash
# Developer tool (8-GPU box), round 2: weak-scaling bench line at N = 8 and BASELINE config 5 as an actual job --
# a 100-hour synthetic 48 kHz corpus through the file-level driver on 8 GPUs (reader / GPU / writer pipeline,
# float32 lf0 / mgc / bap files on disk, one NCCL all-reduce of the statistics partials at the end).
mkdir -p gpurun_out
N=${N:-8}
df -h /tmp | tail -1
nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N exit $?"; tail -2 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_n$N.json'))
print("N=%d value %.0f e2e %.0f ms/step %.1f e2e ms %.1f" % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
PY
rm -rf /tmp/cfg5 && mkdir -p /tmp/cfg5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 hts-train-world_b200/driver.py --synthetic-hours ${HOURS:-100} --out-dir /tmp/cfg5 --fs 48000 > gpurun_out/r2_config5_n$N.log 2>&1; echo "config 5 exit $?"; tail -3 gpurun_out/r2_config5_n$N.log
cp /tmp/cfg5/world_b200_stats.json gpurun_out/r2_config5_n${N}_stats.json 2>/dev/null
ls /tmp/cfg5/mgc | wc -l; du -sh /tmp/cfg5 | tail -1
rm -rf /tmp/cfg5 && mkdir -p /tmp/cfg5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 hts-train-world_b200/driver.py --synthetic-hours ${HOURS:-100} --out-dir /tmp/cfg5 --fs 48000 --no-files > gpurun_out/r2_config5_n${N}_nofiles.log 2>&1; echo "config 5 (no files) exit $?"; tail -2 gpurun_out/r2_config5_n${N}_nofiles.log
cp /tmp/cfg5/world_b200_stats.json gpurun_out/r2_config5_n${N}_nofiles_stats.json 2>/dev/null
rm -rf /tmp/cfg5

for p in 1 2 4 8; do WB_E2E_PARTS=$p timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/e2e_$p.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/e2e_$p.json')); e=d['e2e']; print($p, 'resident', round(d['ms_per_step'],1), 'e2e dev', round(e['device_ms_per_step'],1), 'wall', round(e['wall_ms_per_step'],1))"; done

#!/bin/bash
# e2e of config 2 with one, two, three and four contexts (pipelines) on the same GPU
for cfg in "0,0 8" "0,0,0 8" "0,0,0,0 8" "0,0,0 6" "0,0 12"; do
  set -- $cfg
  echo "== FNFT_B200_DEVICES=$1 FNFT_B200_PIPE=$2"
  export FNFT_B200_DEVICES=$1
  FNFT_B200_PIPE=$2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('  value %.0f ms/step %.2f  e2e %.0f ms/step %.2f %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['ms_each_step_rank0']))
"
done
unset FNFT_B200_DEVICES
echo "== config 4 (kdvv B = 2048), auto contexts from 2048 signals"
FNFT_B200_CTX_MIN_BATCH=2048 python scripts/bench_configs.py --configs 4 --ref-signals 1 2>/dev/null | cut -c1-300
echo "== config 4, default"
python scripts/bench_configs.py --configs 4 --ref-signals 1 2>/dev/null | cut -c1-300

#!/bin/bash
# e2e of config 2 with one and with two contexts (pipelines) on the same GPU
for cfg in "- 8" "0,0 8" "0,0 4" "- 4" "0,0 6"; do
  set -- $cfg
  echo "== FNFT_B200_DEVICES=$1 FNFT_B200_PIPE=$2"
  if [ "$1" = "-" ]; then unset FNFT_B200_DEVICES; else export FNFT_B200_DEVICES=$1; fi
  FNFT_B200_PIPE=$2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('  value %.0f ms/step %.2f  e2e %.0f ms/step %.2f %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['ms_each_step_rank0']))
"
done

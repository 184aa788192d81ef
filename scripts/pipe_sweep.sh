#!/bin/bash
# host-buffer path: number of pipeline chunks / taper
for cfg in "8 1" "4 1" "6 1" "12 1" "16 1" "8 0" "16 0" "32 0"; do
  set -- $cfg
  echo "== FNFT_B200_PIPE=$1 FNFT_B200_PIPE_TAPER=$2"
  FNFT_B200_PIPE=$1 FNFT_B200_PIPE_TAPER=$2 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('  value %.0f ms/step %.2f  e2e %.0f ms/step %.2f %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['ms_each_step_rank0']))
"
done

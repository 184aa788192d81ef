#!/bin/bash
# host-buffer path (two pipelines per GPU, the default): chunks per pipeline / taper depth
for cfg in "8 2" "8 3" "6 3" "5 2" "4 2" "6 2" "4 3"; do
  set -- $cfg
  echo "== FNFT_B200_PIPE=$1 FNFT_B200_PIPE_TAPER=$2"
  FNFT_B200_PIPE=$1 FNFT_B200_PIPE_TAPER=$2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('  value %.0f ms/step %.2f  e2e %.0f ms/step %.2f %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['ms_each_step_rank0']))
"
done

#!/bin/bash
# several contexts on ONE GPU behind one fnft_nsev_batch call (FNFT_B200_DEVICES with a repeated id): their pipelines
# interleave, so kernel tails and the exposed first / last copies of one overlap the kernels of the other
for devs in "" "0,0" "0,0,0" "0,0,0,0"; do
  for pipe in 8 4; do
    echo "== FNFT_B200_DEVICES=$devs FNFT_B200_PIPE=$pipe"
    FNFT_B200_DEVICES=$devs FNFT_B200_PIPE=$pipe python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('  value %.0f ms/step %.2f  e2e %.0f ms/step %.2f %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['ms_each_step_rank0']))
"
  done
done

#!/bin/bash
# cluster variants of the fused tree levels
run() {
  echo "== $*"
  env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']
        print('  value %.0f ms/step %.2f'%(d['value'],d['ms_per_step']), {a.replace('tree_up_smem_',''):round(b,2) for a,b in k.items() if 'smem' in a})
"
}
run A=0
run FNFT_B200_UP12_CLUSTER=1
run FNFT_B200_UP13_L2H=11
run FNFT_B200_UP12_CLUSTER=1 FNFT_B200_UP13_L2H=11

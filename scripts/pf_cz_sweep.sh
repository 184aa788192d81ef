#!/bin/bash
for pf in ${PFS:-0 3 6 12 24}; do
  echo "== FNFT_B200_PFD_CZ=$pf"
  FNFT_B200_PFD_CZ=$pf python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']
        print('  value %.0f ms/step %.2f tree_ms %.2f'%(d['value'],d['ms_per_step'],d['roofline']['tree_ms_per_step']), {a:round(b,2) for a,b in k.items() if 'cz_' in a})
"
done

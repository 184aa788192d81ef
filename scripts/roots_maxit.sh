#!/bin/bash
for it in 200 100 60 40 30; do
  echo "== FNFT_B200_ROOTS_MAXIT=$it"
  FNFT_B200_ROOTS_MAXIT=$it python scripts/bench_configs.py --configs 7 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  %.0f signals/s  %.1f ms  mean_K %.4f all_found %.4f'%(d['value'],d['ms_per_call'],d['mean_K'],d['all_eigenvalues_found']))
"
done

#!/bin/bash
# root finder A/B: kernel with compaction of the moving roots (default) against the one-bit-per-root kernel
for cmp in 0 1; do
  echo "== FNFT_B200_ROOTS_COMPACT=$cmp"
  FNFT_B200_ROOTS_COMPACT=$cmp python scripts/cfg_profile.py 7 1.0 2>&1 | grep -E "poly_roots|total kernel"
  FNFT_B200_ROOTS_COMPACT=$cmp python scripts/bench_configs.py --configs 7 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  %.0f signals/s  %.1f ms  mean_K %.4f all_found %.4f'%(d['value'],d['ms_per_call'],d['mean_K'],d['all_eigenvalues_found']))
"
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "roots or default or fasteig or nsep" 2>&1 | tail -3
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "7 or default" 2>&1 | tail -3

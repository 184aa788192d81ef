#!/bin/bash
# L2 prefetch distance sweep, one value per kernel family and run (FNFT_B200_PFD_<family>, CTAs; 0 = off)
S11=(${S11:-0 4 8 13 20 32}); S12=(${S12:-0 2 4 6 10 16}); S13=(${S13:-0 1 2 4 8 16})
RA=(${RA:-0 2 6 12 24 48}); CO=(${CO:-0 3 6 12 24 48}); RC=(${RC:-0 3 6 12 24 48})
for i in 0 1 2 3 4 5; do
  echo "== smem11=${S11[$i]} smem12=${S12[$i]} smem13=${S13[$i]} rows_a=${RA[$i]} cols=${CO[$i]} rows_c=${RC[$i]}"
  FNFT_B200_PFD_SMEM11=${S11[$i]} FNFT_B200_PFD_SMEM12=${S12[$i]} FNFT_B200_PFD_SMEM13=${S13[$i]} \
  FNFT_B200_PFD_ROWS_A=${RA[$i]} FNFT_B200_PFD_COLS=${CO[$i]} FNFT_B200_PFD_ROWS_C=${RC[$i]} \
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']
        print('  value %.0f ms/step %.2f tree_ms %.2f'%(d['value'],d['ms_per_step'],d['roofline']['tree_ms_per_step']), {a.replace('tree_',''):round(b,2) for a,b in k.items() if 'up_' in a})
"
done

#!/bin/bash
# Does an L2-sized chunk (workspace reused by every chunk, so level buffers stay in the 126 MB L2) beat
# one large chunk?  Device-resident bench at several workspace limits.
for mb in 0 4096 1024 512 384 256 192 128 96; do
  echo "== FNFT_B200_WORKSPACE_MB=$mb"
  FNFT_B200_WORKSPACE_MB=$mb python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']
        print('  value %.0f e2e %.0f ms/step %.2f tree_ms %.1f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['tree_ms_per_step'],d['gpu_launches']), {a.replace('tree_',''):round(b,2) for a,b in k.items()})
"
done

#!/bin/bash
run() {
  echo "== $*"
  env "$@" python bench.py --batch 1024 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']
        print('  value %.0f e2e %.0f tree_ms %.1f'%(d['value'],d['e2e']['value'],d['roofline']['tree_ms_per_step']), {a.replace('tree_pair_',''):round(b,1) for a,b in k.items()})
"
}
run A=1
run A=2

#!/bin/bash
# Runs the config-2 kernel timing (bench.py device-resident leg with per-launch CUDA events) for the main
# library and every variant in fnft_b200/lib/var/ (scripts/variant.sh); one line per library.
#   scripts/kvariants.sh [pattern]      on the GPU box, e.g. through gpurun
cd "$(dirname "$0")/.."
pat=${1:-}
for lib in fnft_b200/lib/libfnft_b200.so fnft_b200/lib/var/*${pat}*.so; do
  [ -f "$lib" ] || continue
  FNFT_B200_LIB=$PWD/$lib python bench.py --no-extras --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=l['roofline']['kernel_ms_per_step']
print('%-28s %8.0f sig/s  tree %.2f ms | %s' % ('$(basename $lib .so)'.replace('libfnft_b200',''), l['value'], l['roofline']['tree_ms_per_step'], ' '.join('%s=%.2f'%(a.replace('tree_','').replace('up_smem_',''),b) for a,b in k.items())))
"
done

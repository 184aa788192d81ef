#!/bin/bash
for nt in 256 512; do
  echo "== NT=$nt"
  FNFT_B200_ROOTS_NT=$nt python scripts/cfg_profile.py 7 1.0 2>&1 | grep -E "poly_roots|total"
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roots" 2>&1 | tail -3
python scripts/bench_configs.py --configs 7 2>/dev/null | grep "^{" | cut -c1-400
CMD="python scripts/cfg_profile.py 7 1.0"
ncu --set full --clock-control none --import-source on -k regex:"k_roots_aberth" -c 1 -f -o gpurun_out/r6d_roots $CMD > gpurun_out/r6d_ncu.log 2>&1
python scripts/ncu_keys.py gpurun_out/r6d_roots.ncu-rep

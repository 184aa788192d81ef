#!/bin/bash
# root finder with compaction: coefficient placement and CTA size, then ncu --set full of the default variant
for cs in 0 1; do for nt in 256 512 1024; do
  echo "== SMEM_COEF=$cs NT=$nt"
  FNFT_B200_ROOTS_SMEM_COEF=$cs FNFT_B200_ROOTS_NT=$nt python scripts/cfg_profile.py 7 1.0 2>&1 | grep -E "poly_roots"
done; done
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roots" 2>&1 | tail -3
CMD="python scripts/cfg_profile.py 7 0.5"
ncu --set full --clock-control none --import-source on -k regex:"k_roots_aberth" -c 1 -f -o gpurun_out/r6c_roots $CMD > gpurun_out/r6c_ncu.log 2>&1
python scripts/ncu_keys.py gpurun_out/r6c_roots.ncu-rep

#!/bin/bash
# root finder: config 7 (fnft_nsev default options on config 3's signals) per-kernel times, and the root-finder tests
for split in 0 1; do
  echo "== FNFT_B200_ROOTS_SPLIT32=$split"
  FNFT_B200_ROOTS_SPLIT32=$split python scripts/cfg_profile.py 7 0.5 2>&1 | grep -E "poly_roots|total kernel|signals/s|wall"
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "roots or default or fasteig or nsep" 2>&1 | tail -3
python -m pytest tests/test_reference_suite.py tests/test_reference_programs.py -m gpu -q 2>&1 | tail -3

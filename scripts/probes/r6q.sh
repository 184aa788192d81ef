#!/bin/bash
python scripts/cfg_profile.py 7 1.0 2>&1 | grep -E "poly_roots|total"
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roots or default or fasteig or nsep" 2>&1 | tail -2
python -m pytest tests/test_gpu_fullsize.py tests/test_reference_programs.py -m gpu -q 2>&1 | tail -2

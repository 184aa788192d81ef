python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roots" 2>&1 | grep -v "^$" | tail -40
FNFT_B200_ROOTS_STATS=1 python scripts/cfg_profile.py 7 1.0 2>&1 | grep -E "roots|total kernel"

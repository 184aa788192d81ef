#!/bin/bash
python scripts/bench_configs.py --configs 7 2>/dev/null | grep "^{" | cut -c1-500
FNFT_B200_NSEV_TIMING=1 python scripts/cfg_profile.py 7 1.0 2>&1 | tail -40

#!/bin/bash
python bench.py > gpurun_out/r6g_bench.json 2> gpurun_out/r6g_bench.err; echo "bench rc $?"
CMD="python scripts/cfg_profile.py 3 1.0"
ncu --set full --clock-control none --import-source on -k regex:"k_newton_warp|k_normconsts_warp" -c 2 -f -o gpurun_out/r6g_bound $CMD > gpurun_out/r6g_ncu.log 2>&1
python scripts/ncu_keys.py gpurun_out/r6g_bound.ncu-rep

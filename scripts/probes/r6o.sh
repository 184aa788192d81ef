#!/bin/bash
# final code: ncu --set full of the Newton / norming-constant kernels (config 3) and of the root finder (config 7)
ncu --set full --clock-control none --import-source on -k regex:"k_newton_warp|k_normconsts_warp" -c 2 -f -o gpurun_out/r6o_bound python scripts/cfg_profile.py 3 1.0 > gpurun_out/r6o_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_roots_aberth" -c 1 -f -o gpurun_out/r6o_roots python scripts/cfg_profile.py 7 1.0 > gpurun_out/r6o_ncu7.log 2>&1
python scripts/ncu_keys.py gpurun_out/r6o_bound.ncu-rep
python scripts/ncu_sass_mix.py gpurun_out/r6o_bound.ncu-rep k_newton_warp | head -3
python scripts/ncu_sass_mix.py gpurun_out/r6o_bound.ncu-rep k_normconsts_warp | head -3
python scripts/ncu_sass_mix.py gpurun_out/r6o_roots.ncu-rep k_roots | head -3

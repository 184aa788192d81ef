#!/bin/bash
# ncu --set full of the general chirp-z column kernels as config 5 (scaled to 128 signals) runs them
TAG=${1:-r4}
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"blk_cz_cols" --launch-skip 4 -c 4 -f -o gpurun_out/${TAG}_nsep \
    python scripts/cfg_profile.py 5 0.125 > gpurun_out/${TAG}_nsep_ncu.log 2>&1
python scripts/ncu_keys.py gpurun_out/${TAG}_nsep.ncu-rep > gpurun_out/${TAG}_nsep_keys.txt 2>&1
tail -2 gpurun_out/${TAG}_nsep_ncu.log

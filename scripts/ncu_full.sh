#!/bin/bash
# `ncu --set full` of every bench-path kernel of one config-2 step at 256 signals per launch -> gpurun_out/<tag>_full.ncu-rep
# and the key figures (pipes, occupancy, stalls) as text (scripts/ncu_keys.py) -> gpurun_out/<tag>_full_keys.txt
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 1 --batch 256 --no-cpu-baseline --no-extras --parity-signals 4"
$CMD > gpurun_out/${TAG}_full_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"k_tree_low2|k_up_|k_cz2_" -c 10 -f -o gpurun_out/${TAG}_full \
    $CMD > gpurun_out/${TAG}_full_ncu.log 2>&1
python scripts/ncu_keys.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_full_keys.txt 2>&1
tail -2 gpurun_out/${TAG}_full_ncu.log

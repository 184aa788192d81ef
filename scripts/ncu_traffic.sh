#!/bin/bash
# DRAM bytes and duration of every bench-path kernel launch of one config-2 step at 1024 signals per launch
# (ncu replays each kernel, so the bench numbers printed under ncu are meaningless); -> gpurun_out/<tag>_traffic.csv
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline --no-extras --parity-signals 4"
$CMD > gpurun_out/${TAG}_traffic_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
    -k regex:"k_tree_low2|k_up_|k_cz2_" -c 200 --csv --log-file gpurun_out/${TAG}_traffic.csv $CMD > gpurun_out/${TAG}_traffic_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_traffic_ncu.log

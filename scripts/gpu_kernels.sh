#!/bin/bash
# Per-kernel table visit: ncu time / DRAM bytes / instruction counts of every kernel around the headline scan.
# usage (under gpurun): bash scripts/gpu_kernels.sh <tag>
tag=${1:-kern}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --csv --log-file $out/kernels_launches.csv python scripts/kernels_target.py > $out/kernels.log 2>&1
python scripts/kernels_table.py $out/kernels_launches.csv $out/kernels_table.md | tail -30

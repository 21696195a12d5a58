#!/bin/bash
# ncu check of the L2-residency variants (scratch).  usage: bash scripts/gpu_pin_ncu.sh <tag>
tag=${1:-pinncu}
out=gpurun_out/$tag
mkdir -p $out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 ncu --clock-control none --cache-control none --metrics $M -k regex:two_opt_scan_matrix -s 6 -c 3 --csv --log-file $out/ncu_$name.csv python scripts/prof_target.py matrix 10000 12 best nint > $out/ncu_$name.log 2>&1
  grep -E "two_opt_scan" $out/ncu_$name.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | paste - - - - - - | head -3 | sed "s/^/$name: /"
}
run pin0 TL_MAT_PIN_MB=0
run hint70 TL_MAT_PIN_MB=70
run win60 TL_MAT_PIN_MB=60 TL_MAT_PIN_MODE=window
timeout 600 python scripts/perf_probe.py matrix:5000 matrix:10000 matrix:20000 nint:14000 2>&1 | tee $out/probe_sizes.txt

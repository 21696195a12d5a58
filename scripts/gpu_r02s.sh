#!/bin/bash
out=gpurun_out/r02s
mkdir -p $out
for cfg in 3 4; do
echo "== cluster timing, cluster CTA cfg $cfg"; TL_BATCH_CLUSTER_CFG=$cfg timeout 120 python scripts/batch_cluster_timing.py 2,4,8 2>&1 | grep -E "B512|B256|B128|B64|B16" | tee $out/batch_cluster_timing_cfg$cfg.txt
done

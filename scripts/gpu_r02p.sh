#!/bin/bash
out=gpurun_out/r02p
mkdir -p $out
echo "== pytest batch"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > $out/pytest_batch.txt 2>&1; tail -3 $out/pytest_batch.txt
echo "== cluster timing"; timeout 900 python scripts/batch_cluster_timing.py 1,2,4,8 $out/batch_cluster_timing.json 2>&1 | tee $out/batch_cluster_timing.txt
for ipw in 2 5; do
echo "== cluster timing IPW=$ipw"; TL_BATCH_IPW=$ipw timeout 900 python scripts/batch_cluster_timing.py 1,2,4 2>&1 | grep -E "B128|B256|B512" | tee $out/batch_cluster_timing_ipw$ipw.txt
done

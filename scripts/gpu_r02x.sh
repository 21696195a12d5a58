#!/bin/bash
out=gpurun_out/r02x
mkdir -p $out
echo "== pytest batch"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > $out/pytest_batch.txt 2>&1; tail -3 $out/pytest_batch.txt
echo "== steal on"; timeout 120 python scripts/batch_cluster_timing.py auto 2>&1 | tee $out/batch_steal_on.txt
echo "== steal off"; TL_BATCH_STEAL=0 timeout 120 python scripts/batch_cluster_timing.py auto 2>&1 | tee $out/batch_steal_off.txt

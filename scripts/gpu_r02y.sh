#!/bin/bash
out=gpurun_out/r02y
mkdir -p $out
echo "== pytest batch"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > $out/pytest_batch.txt 2>&1; tail -3 $out/pytest_batch.txt
echo "== timing"; timeout 120 python scripts/batch_cluster_timing.py auto,1 2>&1 | tee $out/batch_timing.txt

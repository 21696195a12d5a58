#!/bin/bash
out=gpurun_out/r02o
mkdir -p $out
echo "== pytest batch"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > $out/pytest_batch.txt 2>&1; tail -15 $out/pytest_batch.txt
echo "== cluster timing"; timeout 900 python scripts/batch_cluster_timing.py 1,2,4,8,auto $out/batch_cluster_timing.json 2>&1 | tee $out/batch_cluster_timing.txt
echo "== host mirror"; timeout 900 python -m pytest tests/test_host_mirror.py -m gpu -x -q 2>&1 | tail -5

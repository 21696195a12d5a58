#!/bin/bash
out=gpurun_out/r02t
mkdir -p $out
echo "== pytest batch"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > $out/pytest_batch.txt 2>&1; tail -3 $out/pytest_batch.txt
echo "== cluster timing (auto policy)"; timeout 120 python scripts/batch_cluster_timing.py auto $out/batch_auto_timing.json 2>&1 | tee $out/batch_auto_timing.txt
echo "== full pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -5 $out/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 3000 $out/bench.json

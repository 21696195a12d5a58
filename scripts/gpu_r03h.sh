#!/bin/bash
out=gpurun_out/r03h
mkdir -p $out
echo "== pytest knn/nn"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn or nn_ or nn or smoke" 2>&1 | tail -3
echo "== nn timing"; timeout 120 python scripts/nn_timing.py 2>&1 | tee $out/nn_timing.txt

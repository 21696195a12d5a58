#!/bin/bash
out=gpurun_out/r03i
mkdir -p $out
echo "== pytest cached"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cached" > $out/pytest_cached.txt 2>&1; tail -25 $out/pytest_cached.txt

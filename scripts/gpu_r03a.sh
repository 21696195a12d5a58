#!/bin/bash
out=gpurun_out/r03a
mkdir -p $out
echo "== pytest ga"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_ga_" 2>&1 | tail -3
echo "== ga timing"; timeout 300 python scripts/ga_timing.py 2>&1 | tee $out/ga_timing.txt

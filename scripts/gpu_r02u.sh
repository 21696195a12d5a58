#!/bin/bash
out=gpurun_out/r02u
mkdir -p $out
echo "== pytest k1"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k1 or nint or matrix" > $out/pytest_k1.txt 2>&1; tail -3 $out/pytest_k1.txt
echo "== k1 timing"; TL_K1_TIMING=1 timeout 120 python scripts/k1_timing.py 2>&1 | tee $out/k1_timing.txt

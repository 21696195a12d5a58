#!/bin/bash
# Mode R persistent kernel with the nint metrics (matrix sessions of coordinate problems)
out=gpurun_out/r03w
mkdir -p $out
TL_REF_CLUSTER=0 timeout 300 python scripts/ref_persist_timing.py perstep 2>&1 | grep -v resumed | grep nint | tee $out/ref_persist_timing.txt
timeout 300 python scripts/ref_persist_timing.py persistent 2>&1 | tee -a $out/ref_persist_timing.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mode_r or golden or smoke or nint or matrix_path or explicit" 2>&1 | tail -5 | tee $out/pytest_mode_r.txt

#!/bin/bash
# Mode R: persistent cluster kernel vs per-step kernel (same moves?), then the Mode R parity tests
out=gpurun_out/r03r
mkdir -p $out
TL_REF_CLUSTER=0 timeout 300 python scripts/ref_persist_timing.py perstep 2>&1 | grep -v resumed | tee -a $out/ref_persist_timing2.txt
TL_NO_SCREEN=1 timeout 300 python scripts/ref_persist_timing.py cluster16-noscreen 2>&1 | tee -a $out/ref_persist_timing2.txt
timeout 300 python scripts/ref_persist_timing.py cluster16-screen 2>&1 | tee -a $out/ref_persist_timing2.txt
TL_REF_CLUSTER=8 timeout 300 python scripts/ref_persist_timing.py cluster8-screen 2>&1 | grep -v resumed | tee -a $out/ref_persist_timing2.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mode_r or golden or smoke" 2>&1 | tail -5 | tee $out/pytest_mode_r.txt

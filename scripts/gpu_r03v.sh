#!/bin/bash
# Mode R persistent kernel, defaults of the round (512 threads, 4-row units, one-round first window): timing + Mode R tests
out=gpurun_out/r03v
mkdir -p $out
timeout 300 python scripts/ref_persist_timing.py default 2>&1 | tee $out/ref_persist_timing.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mode_r or golden or smoke" 2>&1 | tail -5 | tee $out/pytest_mode_r.txt

#!/bin/bash
# Mode R persistent kernel: where do the cycles go (profiling build)
out=gpurun_out/r03t
mkdir -p $out
TL_LIB=variants/lib_refp_prof.so timeout 300 python scripts/ref_persist_timing.py prof 2>&1 | grep -v resumed | tee -a $out/ref_persist_prof.txt

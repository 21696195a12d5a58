#!/bin/bash
out=gpurun_out/r02q
mkdir -p $out
echo "== phases (TL_TIMELINE build)"; TL_LIB=variants/lib_tline.so TL_BATCH_PHASES=1 timeout 900 python scripts/batch_cluster_timing.py 1,2,4 2>&1 | grep -E "B128|B512|B1024|B64|phases" | tee $out/batch_phases.txt

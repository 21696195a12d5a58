#!/bin/bash
# L2-residency sweep of the matrix-path scan (scratch).  usage: bash scripts/gpu_pin.sh <tag>
tag=${1:-pin}
out=gpurun_out/$tag
mkdir -p $out
envs="TL_MAT_PIN_MB=0"
for mb in 30 50 70 90 110; do envs="$envs,TL_MAT_PIN_MB=$mb"; done
for mb in 40 60 80; do envs="$envs,TL_MAT_PIN_MB=$mb;TL_MAT_PIN_MODE=window"; done
envs="$envs,TL_MAT_PIN_MB=100;TL_MAT_PIN_MODE=window;TL_MAT_PIN_HIT=0.6"
PROBE_ENVS="$envs" timeout 900 python scripts/perf_probe.py matrix:10000 nint:10000 2>&1 | tee $out/probe.txt

#!/bin/bash
out=gpurun_out/r02h
mkdir -p $out
echo "== pytest batch"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or config5" 2>&1 | tail -5 | tee $out/pytest.txt
for ch in 64 128; do
  echo "== batch scaling engine=pop chunk=$ch"; TL_POP_CHUNK=$ch TL_BATCH_ENGINE=pop WORLDS=1,4,8 timeout 600 python scripts/batch_scaling.py $out/batch_scaling_pop_chunk$ch.json 2>&1 | tee $out/batch_scaling_pop_chunk$ch.txt
done

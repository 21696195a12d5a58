#!/bin/bash
# round 2: K2-pop with the FIFO scheduler; scalar fixed-column HBM pattern microbench
out=gpurun_out/r02e
mkdir -p $out
echo "== pytest batch"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or config5" 2>&1 | tail -15 | tee $out/pytest_batch.txt
for eng in pop; do
  echo "== batch scaling engine=$eng"; TL_BATCH_ENGINE=$eng timeout 600 python scripts/batch_scaling.py $out/batch_scaling_$eng.json 2>&1 | tee $out/batch_scaling_$eng.txt
done
for ch in 32; do
  echo "== batch scaling engine=pop chunk=$ch"; TL_POP_CHUNK=$ch TL_BATCH_ENGINE=pop WORLDS=1,8 timeout 600 python scripts/batch_scaling.py $out/batch_scaling_pop_chunk$ch.json 2>&1 | tee $out/batch_scaling_pop_chunk$ch.txt
done
echo "== hbm patterns"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_patterns scripts/hbm_patterns.cu && timeout 300 /tmp/hbm_patterns 10000 2>&1 | head -24 | tee $out/hbm_patterns.txt

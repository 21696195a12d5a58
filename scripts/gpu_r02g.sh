#!/bin/bash
out=gpurun_out/r02g
mkdir -p $out
echo "== pytest aco+batch"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aco or batch or config5" 2>&1 | tail -15 | tee $out/pytest.txt
echo "== aco timing"; timeout 600 python scripts/aco_timing.py 2>&1 | tee $out/aco_timing.txt
for ch in 64 128 256; do
  echo "== batch scaling engine=pop chunk=$ch"; TL_POP_CHUNK=$ch TL_BATCH_ENGINE=pop WORLDS=1,4,8 timeout 600 python scripts/batch_scaling.py $out/batch_scaling_pop_chunk$ch.json 2>&1 | tee $out/batch_scaling_pop_chunk$ch.txt
done

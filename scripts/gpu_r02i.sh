#!/bin/bash
out=gpurun_out/r02i
mkdir -p $out
echo "== pytest batch"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or config5" 2>&1 | tail -5 | tee $out/pytest.txt
echo "== default"; TL_BATCH_ENGINE=pop WORLDS=1,2,4,8 timeout 600 python scripts/batch_scaling.py $out/bs_default.json 2>&1 | tee $out/bs_default.txt
for g in 1 2 4; do
  echo "== group=$g"; TL_POP_GROUP=$g TL_BATCH_ENGINE=pop WORLDS=1,8 timeout 600 python scripts/batch_scaling.py $out/bs_g$g.json 2>&1 | tee $out/bs_g$g.txt
done
for ch in 32 64 96; do
  echo "== chunk=$ch"; TL_POP_CHUNK=$ch TL_BATCH_ENGINE=pop WORLDS=1,8 timeout 600 python scripts/batch_scaling.py $out/bs_c$ch.json 2>&1 | tee $out/bs_c$ch.txt
done

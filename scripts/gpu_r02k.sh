#!/bin/bash
out=gpurun_out/r02k
mkdir -p $out
echo "== pytest batch"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or config5" 2>&1 | tail -5 | tee $out/pytest.txt
for e in pop cta; do TL_BATCH_ENGINE=$e timeout 300 python scripts/batch_phases.py | tee -a $out/phases.txt; done
for e in pop cta; do B=128 TL_BATCH_ENGINE=$e timeout 300 python scripts/batch_phases.py | tee -a $out/phases.txt; done
TL_POP_CHUNK=96 TL_BATCH_ENGINE=pop timeout 300 python scripts/batch_phases.py | tee -a $out/phases.txt
B=128 TL_POP_CHUNK=96 TL_BATCH_ENGINE=pop timeout 300 python scripts/batch_phases.py | tee -a $out/phases.txt

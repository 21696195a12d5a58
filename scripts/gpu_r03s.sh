#!/bin/bash
# Mode R persistent kernel: cluster-wide early exit; variants of (threads per CTA, rows per unit)
out=gpurun_out/r03s
mkdir -p $out
timeout 300 python scripts/ref_persist_timing.py t512-rb4 2>&1 | tee -a $out/ref_persist_timing2.txt
TL_REF_CLUSTER=8 timeout 300 python scripts/ref_persist_timing.py t512-rb4-cluster8 2>&1 | grep -v resumed | tee -a $out/ref_persist_timing2.txt
for v in refp_t1024_rb4 refp_t512_rb8 refp_t256_rb4; do
  TL_LIB=variants/lib_$v.so timeout 300 python scripts/ref_persist_timing.py $v 2>&1 | grep -v resumed | tee -a $out/ref_persist_timing2.txt
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mode_r or golden or smoke" 2>&1 | tail -5 | tee $out/pytest_mode_r.txt

#!/bin/bash
# Mode R persistent kernel: shared-memory arrays instead of 16-byte records; window policy variants
out=gpurun_out/r03u
mkdir -p $out
timeout 300 python scripts/ref_persist_timing.py soa 2>&1 | tee -a $out/ref_persist_timing.txt
for v in refp_prof refp_r1 refp_r3 refp_g2 refp_g8; do
  TL_LIB=variants/lib_$v.so timeout 300 python scripts/ref_persist_timing.py $v 2>&1 | grep -v resumed | tee -a $out/ref_persist_timing.txt
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mode_r or golden or smoke" 2>&1 | tail -5 | tee $out/pytest_mode_r.txt

#!/bin/bash
# Mode R persistent kernel, lean loop: rows per unit / threads per CTA variants
out=gpurun_out/r03z
mkdir -p $out
for v in refp_rb8 refp_rb6 refp_rb2 refp_t768; do
  TL_LIB=variants/lib_$v.so timeout 300 python scripts/ref_persist_timing.py $v 2>&1 | grep -v resumed | grep "n=1000 f32\|n=10000\|n=14000" | tee -a $out/ref_persist_variants.txt
done

#!/bin/bash
out=gpurun_out/r03f
mkdir -p $out
timeout 100 python scripts/scan_variant_timing.py 2>&1 | grep "^default" | tee $out/scan_variants.txt
for v in a b c d e; do TL_LIB=variants/lib_scan_$v.so timeout 100 python scripts/scan_variant_timing.py 2>&1 | grep "^variants" | tee -a $out/scan_variants.txt; done

#!/bin/bash
# compute-sanitizer over the persistent Mode R kernel (memcheck, racecheck on shared memory, synccheck)
out=gpurun_out/r04e
mkdir -p $out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 200 compute-sanitizer --tool $tool --kernel-name kernel_substring=ref_persistent python scripts/refp_sanitize.py > $out/sanitize_$tool.txt 2>&1
  tail -4 $out/sanitize_$tool.txt
done

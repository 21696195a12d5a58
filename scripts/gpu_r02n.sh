#!/bin/bash
out=gpurun_out/r02n
mkdir -p $out
for r in 1 2 3 4 5; do
echo "== pytest ga rep $r"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ga_ or aco" > $out/pytest_rep$r.txt 2>&1; echo "rc=$?"; grep -E "passed|failed|Fatal|File \"/root" $out/pytest_rep$r.txt | head -20
done

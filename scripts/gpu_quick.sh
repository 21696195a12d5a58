#!/bin/bash
# Quick GPU-box visit: parity tests + perf probe (+ optional bench).  usage: bash scripts/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -15 | tee $out/pytest.txt
echo "== probe"; timeout 600 python scripts/perf_probe.py 1000 10000 100000 2>&1 | tee $out/probe.txt
echo "== bench"; timeout 600 python bench.py 2>$out/bench.err | tee $out/bench.json
tail -3 $out/bench.err

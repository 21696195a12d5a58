#!/bin/bash
# round 2 first visit: parity tests (incl. BASELINE-size ones), smoke, bench at the driver's --steps 20 and at the default
out=gpurun_out/r02a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
nproc > $out/nproc.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 | tee $out/pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $out/smoke.txt
echo "== bench steps20"; timeout 600 python bench.py --steps 20 --warmup 3 2>$out/bench20.err | tee $out/bench20.json
echo "== bench default"; timeout 600 python bench.py 2>$out/bench.err | tee $out/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>$out/bench_reference.err | tee $out/bench_reference.json

#!/bin/bash
out=gpurun_out/r02m
mkdir -p $out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $out/pytest.txt
echo "== k1 timing"; TL_K1_TIMING=1 timeout 300 python scripts/k1_timing.py 2>&1 | tee $out/k1_timing.txt
echo "== session create timing (k1_square inside)"; TL_DEBUG_TIMING=1 timeout 600 python bench.py --steps 20 --no-partitioned 2>$out/bench.err > $out/bench.json; grep -E "\[tl\]|\[bench\]" $out/bench.err | tail -6; cat $out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['wall_ms_per_call_all'],'frac',d['roofline']['frac'])"

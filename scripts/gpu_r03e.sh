#!/bin/bash
out=gpurun_out/r03e
mkdir -p $out
echo "== bench --steps 20"; timeout 600 python bench.py --steps 20 --warmup 3 --no-partitioned > $out/bench_steps20.json 2> $out/bench_steps20.err; tail -3 $out/bench_steps20.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03e/bench_steps20.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
print(json.dumps(d['wall_to_local_optimum'],indent=1))
PY

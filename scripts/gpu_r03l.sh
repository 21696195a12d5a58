#!/bin/bash
out=gpurun_out/r03l
mkdir -p $out
echo "== full pytest"; timeout 1800 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -4 $out/pytest_gpu.txt
echo "== bench --steps 20"; timeout 600 python bench.py --steps 20 --warmup 3 --no-partitioned > $out/bench_steps20.json 2> $out/bench_steps20.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03l/bench_steps20.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],'e2e',d['e2e']['value'])
w=d['wall_to_local_optimum']
for k,v in w.items():
    if isinstance(v,dict): print(k, v['wall_ms'], v['device_ms'], v['moves'], v.get('evals'))
print({k:(v['ms_per_step'],v['kernel_ms']) for k,v in d['other_paths'].items()})
PY
echo "== stress"; bash scripts/gpu_r03k.sh

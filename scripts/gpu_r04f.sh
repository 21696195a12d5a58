#!/bin/bash
# reversal helper: thread 0 reads the first entering edge up front (no thread touches a field its neighbour rewrites);
# full parity suite, racecheck of the persistent Mode R kernel, headline bench
out=gpurun_out/r04f
mkdir -p $out
echo "== full pytest"; timeout 1800 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -4 $out/pytest_gpu.txt
echo "== racecheck"; timeout 200 compute-sanitizer --tool racecheck --kernel-name kernel_substring=ref_persistent python scripts/refp_sanitize.py > $out/sanitize_racecheck.txt 2>&1; grep -v "Host Frame\|host backtrace" $out/sanitize_racecheck.txt | tail -6 | cut -c1-250
echo "== bench --steps 20"; timeout 300 python bench.py --steps 20 --warmup 3 --no-partitioned > $out/bench_steps20.json 2> $out/bench_steps20.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04f/bench_steps20.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],[round(x,1) for x in d['e2e']['wall_ms_per_call_all']])
for k,v in d['wall_to_local_optimum'].items():
    if isinstance(v,dict): print(k, round(v['wall_ms'],2), round(v['device_ms'],2), v['moves'])
PY

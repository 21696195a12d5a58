#!/bin/bash
# where do the 10-50 ms outliers of the end-to-end nint matrix calls come from?
out=gpurun_out/r03q
mkdir -p $out
TL_DEBUG_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/trace.txt
grep -c . $out/trace.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03q/bench.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['value'], [round(x,1) for x in d['e2e']['wall_ms_per_call_all']])
for k,v in d['wall_to_local_optimum'].items():
    if isinstance(v,dict): print(k, round(v['wall_ms'],2), round(v['device_ms'],2), [round(x,1) for x in v.get('wall_ms_all',[])])
PY

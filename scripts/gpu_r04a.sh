#!/bin/bash
# validation of the tree with the persistent Mode R kernel + ncu evidence for it
out=gpurun_out/r04a
mkdir -p $out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $out/smoke.txt
echo "== full pytest"; timeout 1800 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -4 $out/pytest_gpu.txt
echo "== bench --steps 20"; timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_steps20.json 2> $out/bench_steps20.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04a/bench_steps20.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],'e2e',d['e2e']['value'],[round(x,1) for x in d['e2e']['wall_ms_per_call_all']],'cpu',d['cpu_baseline']['value'])
for k,v in d['wall_to_local_optimum'].items():
    if isinstance(v,dict): print(k, round(v['wall_ms'],2), round(v['device_ms'],2), v['moves'], v.get('evals'), [round(x,1) for x in v.get('wall_ms_all',[])])
print('config5', d['partitioned']['config5_1024_tours']['single_gpu'])
PY
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-200 $out/bench_reference.json
echo "== ncu launch list, Mode R 10k f32"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $out/launches_mode_r.csv python scripts/prof_target.py recompute 10000 1000000 ref > $out/lt1.log 2>&1; tail -1 $out/lt1.log
echo "== ncu full capture, persistent Mode R kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ref_persistent -c 1 -o $out/ncu_full_mode_r_persistent -f python scripts/prof_target.py recompute 10000 1000000 ref > $out/lt2.log 2>&1; tail -1 $out/lt2.log
ls -la $out

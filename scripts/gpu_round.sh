#!/bin/bash
# One GPU-box visit: parity tests, bench (both paths), ncu launch lists and full captures.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
echo "== bench recompute"; timeout 600 python bench.py --path recompute 2>$out/bench_recompute.err | tee $out/bench_recompute.json
echo "== bench matrix";   timeout 600 python bench.py --path matrix 2>$out/bench_matrix.err | tee $out/bench_matrix.json
echo "== probe"; timeout 600 python scripts/perf_probe.py 1000 10000 100000 2>&1 | tee $out/probe.txt
NCU="ncu --clock-control none"
for p in recompute matrix; do
  timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $out/launches_${p}.csv python scripts/prof_target.py $p 10000 20 > $out/launches_${p}.log 2>&1
  timeout 900 $NCU --set full --import-source on -k regex:two_opt_scan -s 3 -c 2 -f -o $out/prof_${p} python scripts/prof_target.py $p 10000 8 > $out/prof_${p}.log 2>&1
done
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $out/launches_oropt.csv python scripts/prof_target.py recompute 10000 10 oropt > $out/launches_oropt.log 2>&1
ls -la $out

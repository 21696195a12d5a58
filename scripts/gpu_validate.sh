#!/bin/bash
# Full validation visit: parity tests, smoke, both bench arms, launch list + full ncu capture of the headline (int32 matrix) scan.
# usage (under gpurun): bash scripts/gpu_validate.sh <tag>
tag=${1:-val}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $out/smoke.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference 2>$out/bench_reference.err | tee $out/bench_reference.json
echo "== bench"; timeout 600 python bench.py 2>$out/bench.err | tee $out/bench.json
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $out/launches_nint.csv python scripts/prof_target.py matrix 10000 20 best nint > $out/launches_nint.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:two_opt_scan -s 3 -c 1 -f -o $out/prof_nint python scripts/prof_target.py matrix 10000 8 best nint > $out/prof_nint.log 2>&1
ls -la $out
# per-kernel table (time / DRAM bytes / instructions of every kernel around the headline scan)
timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --csv --log-file $out/kernels_launches.csv python scripts/kernels_target.py > $out/kernels.log 2>&1
python scripts/kernels_table.py $out/kernels_launches.csv $out/kernels_table.md > /dev/null
# full capture of the Or-opt scan
timeout 900 $NCU --set full --import-source on -k regex:or_opt_scan -s 1 -c 1 -f -o $out/prof_oropt python scripts/prof_target.py recompute 10000 3 oropt > $out/prof_oropt.log 2>&1
ls -la $out

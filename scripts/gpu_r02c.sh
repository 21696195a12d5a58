#!/bin/bash
# round 2, third visit (1 GPU): parity tests after K1/K4/ABI changes, kernel timings, batch scaling, e2e phases
out=gpurun_out/r02c
mkdir -p $out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
echo "== k1 timing"; TL_K1_TIMING=1 timeout 300 python scripts/k1_timing.py 2>&1 | tee $out/k1_timing.txt
echo "== k1 timing f64 path"; TL_NINT_F64=1 TL_K1_TIMING=1 timeout 300 python scripts/k1_timing.py 2>&1 | tee $out/k1_timing_f64.txt
echo "== k4 timing"; TL_K4_TIMING=1 timeout 300 python scripts/k4_timing.py 2>&1 | tee $out/k4_timing.txt
echo "== k4 timing old kernel"; TL_K4_NO_WARP=1 TL_K4_TIMING=1 timeout 300 python scripts/k4_timing.py 2>&1 | tee $out/k4_timing_old.txt
echo "== batch scaling"; timeout 600 python scripts/batch_scaling.py $out/batch_scaling.json 2>&1 | tee $out/batch_scaling.txt
echo "== e2e phases"; TL_DEBUG_TIMING=1 timeout 600 python bench.py --steps 20 --no-partitioned 2>$out/bench_phases.err > $out/bench_phases.json; grep -E "\[tl\]|\[bench\]" $out/bench_phases.err | tail -20

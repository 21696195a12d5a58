#!/bin/bash
out=gpurun_out/r02f
mkdir -p $out
echo "== hbm patterns (bulk)"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_patterns scripts/hbm_patterns.cu && timeout 300 /tmp/hbm_patterns 10000 2>&1 | head -28 | tee $out/hbm_patterns.txt
echo "== pop quick"; TL_BATCH_ENGINE=pop timeout 120 python scripts/pop_prof_target.py 1024 40 | tee $out/pop_quick.txt
TL_BATCH_ENGINE=cta timeout 120 python scripts/pop_prof_target.py 1024 40 | tee -a $out/pop_quick.txt
echo "== ncu pop"; TL_BATCH_ENGINE=pop timeout 600 ncu --set full --import-source on --clock-control none -k regex:two_opt_pop_kernel -c 1 -f -o $out/prof_pop python scripts/pop_prof_target.py 1024 15 > $out/prof_pop.log 2>&1; tail -3 $out/prof_pop.log
echo "== ncu cta"; TL_BATCH_ENGINE=cta timeout 600 ncu --set full --import-source on --clock-control none -k regex:two_opt_batch_kernel -c 1 -f -o $out/prof_cta python scripts/pop_prof_target.py 1024 15 > $out/prof_cta.log 2>&1; tail -3 $out/prof_cta.log
ls -la $out

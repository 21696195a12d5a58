#!/bin/bash
out=gpurun_out/r02j
mkdir -p $out
for e in pop cta; do TL_BATCH_ENGINE=$e timeout 300 python scripts/pop_prof_target2.py 1024 | tee -a $out/quick.txt; done
TL_POP_CHUNK=96 TL_BATCH_ENGINE=pop timeout 300 python scripts/pop_prof_target2.py 1024 | tee -a $out/quick.txt
TL_BATCH_ENGINE=pop timeout 600 ncu --set full --import-source on --clock-control none -k regex:two_opt_pop_kernel -c 1 -f -o $out/prof_pop python scripts/pop_prof_target2.py 1024 > $out/prof_pop.log 2>&1; tail -2 $out/prof_pop.log
TL_BATCH_ENGINE=cta timeout 600 ncu --set full --import-source on --clock-control none -k regex:two_opt_batch_kernel -s 1 -c 1 -f -o $out/prof_cta python scripts/pop_prof_target2.py 1024 > $out/prof_cta.log 2>&1; tail -2 $out/prof_cta.log

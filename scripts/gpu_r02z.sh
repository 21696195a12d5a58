#!/bin/bash
# ncu evidence for round 2: headline matrix scan (launch list + one full capture), K2-batch, K8
out=gpurun_out/r02z
mkdir -p $out
echo "== launch list, headline path"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_matrix_nint.csv python scripts/prof_target.py matrix 10000 40 best nint > $out/lt1.log 2>&1; tail -2 $out/lt1.log
echo "== full capture, headline scan"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:two_opt_scan_matrix -s 20 -c 1 -o $out/ncu_full_scan_matrix_nint -f python scripts/prof_target.py matrix 10000 40 best nint > $out/lt2.log 2>&1; tail -2 $out/lt2.log
echo "== full capture, K2-batch 128 tours x 1024 threads (30 moves per tour)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:two_opt_batch -c 1 -o $out/ncu_full_batch_128 -f python scripts/batch_prof_target.py 128 30 > $out/lt3.log 2>&1; tail -2 $out/lt3.log
echo "== full capture, K2-batch 1024 tours (cluster 2 x 256, 10 moves per tour)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:two_opt_batch -c 1 -o $out/ncu_full_batch_1024 -f python scripts/batch_prof_target.py 1024 10 > $out/lt4.log 2>&1; tail -2 $out/lt4.log
echo "== launch list + full capture, K8 GA n=1000"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $out/launches_ga.csv python scripts/ga_prof_target.py 1000 10 > $out/lt5.log 2>&1; tail -2 $out/lt5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ga_breed -s 3 -c 1 -o $out/ncu_full_ga_breed -f python scripts/ga_prof_target.py 1000 10 > $out/lt6.log 2>&1; tail -2 $out/lt6.log
ls -la $out

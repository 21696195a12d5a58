"""Profiling target for K2-batch (run under ncu): `batch_prof_target.py <batch> <max_moves>` -- one
tl_two_opt_batch call on the 1000-city instance (shuffled start tours)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, teeline_b200 as T
B, mm = int(sys.argv[1]), int(sys.argv[2])
ctx = T.Context(0)
x, y = bench.gen_uniform(1000, 1000)
p = T.Problem.euc2d(ctx, x, y)
tours = bench.shuffle_tours(1000, range(1, B + 1))
got, st, lengths = p.two_opt_batch(tours, max_moves=mm)
print({"moves": int(st.moves), "scans": int(st.passes), "device_ms": st.device_ms})

"""Time of the batched engines by phase of the search: 1024 random start tours, max_moves = 50..all."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, teeline_b200 as T
n, B = 1000, int(os.environ.get("B", "1024"))
x, y = bench.gen_uniform(n, n)
ctx = T.Context(0)
p = T.Problem.euc2d(ctx, x, y)
tours = bench.shuffle_tours(n, range(1, B + 1))
p.two_opt_batch(tours[:8], max_moves=2)
for mm in (50, 200, 500, 900, -1):
    _, st, _ = p.two_opt_batch(tours, max_moves=mm)
    print(os.environ.get("TL_BATCH_ENGINE"), "B", B, "max_moves", mm, "device_ms %.2f" % st.device_ms, "moves", int(st.moves), flush=True)

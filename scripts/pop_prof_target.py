"""ncu target: one launch of the K2-pop engine (1024 x 1000-city tours, a few moves each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, teeline_b200 as T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 15
n = 1000
x, y = bench.gen_uniform(n, n)
ctx = T.Context(0)
p = T.Problem.euc2d(ctx, x, y)
tours = bench.shuffle_tours(n, range(1, B + 1))
got, st, lengths = p.two_opt_batch(tours, max_moves=moves)
print("moves", int(st.moves), "device_ms", st.device_ms, "evals/s", int(st.evals) / (st.device_ms * 1e-3))

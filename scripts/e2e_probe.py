"""Where does an end-to-end call spend its time?  (GPU box)  usage: e2e_probe.py [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, teeline_b200 as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
torch.cuda.init()
ctx = T.Context(0, stream=torch.cuda.current_stream().cuda_stream)
x, y = bench.gen_grid(n, n)
p0 = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
start = p0.nn_tour(3)
for rep in range(4):
    t0 = time.perf_counter()
    p = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
    t1 = time.perf_counter()
    tour, st, _ = p.local_search(T.ALGO_TWO_OPT_BEST, start, path=T.PATH_MATRIX)
    t2 = time.perf_counter()
    p.close()
    t3 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.2f} ms, local_search {1e3*(t2-t1):.2f} ms (device {st.device_ms:.2f} ms, {int(st.moves)} moves, "
          f"{int(st.launches)} launches), close {1e3*(t3-t2):.2f} ms")

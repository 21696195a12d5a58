"""Profiling target for the kernels around the headline scan (run under ncu on the GPU box): one call
of each entry point at the BASELINE sizes, so that `ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum` gives every kernel's time, DRAM
traffic and instruction count.  scripts/kernels_table.py turns the CSV into the roofline table."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import teeline_b200 as T  # noqa: E402

ctx = T.Context(0)
n = 10000
x, y = bench.gen_uniform(n, n)
p = T.Problem.euc2d(ctx, x, y)
p.matrix_packed()                                   # K1 packed f32, 49 995 000 pairs
gx, gy = bench.gen_grid(n, n)
pi = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
pi.matrix_packed()                                  # K1 packed nint i32
p.knn(5)                                            # K5 k=5 at 10k (LK candidate lists)
start = p.nn_tour(3)                                # K5 k=32 + N1 walk
s = p.session(T.ALGO_OR_OPT, start)                 # K3 at 10k
s.enqueue(2)
ctx.sync()
s.close()
s = p.session(T.ALGO_TWO_OPT_REF, start)            # K2-R at 10k: 40 find/apply pairs
s.enqueue(40)
ctx.sync()
s.close()

n1 = 1000
x1, y1 = bench.gen_uniform(n1, n1)
p1 = T.Problem.euc2d(ctx, x1, y1)
rng = np.random.default_rng(1)
tours = np.stack([rng.permutation(n1) for _ in range(4096)]).astype(np.uint32)
big = np.tile(tours, (16, 1))                       # 65 536 tours x 1000 = 262 MB of indices
p1.tour_lengths(big, T.LEN_EXACT)                   # K4 exact
p1.tour_lengths(big, T.LEN_FAST)                    # K4 fast
p1.tour_lengths(tours[:1024], T.LEN_EXACT)          # config 5 / GA population shape
p1.two_opt_batch(tours[:1024], max_moves=50)        # K2-batch, 1024 x 50 scans (clusters of 2 x 256)
p1.two_opt_batch(tours[:128], max_moves=50)         # K2-batch, 128 tours, one 1024-thread CTA each
p1.two_opt_batch(tours[:16], max_moves=50)          # K2-batch, 16 tours, clusters of 8 x 1024
p1.aco(1, init_tour=tours[0], epochs=2, num_ants=148)  # K7: 2 epochs x 148 ants
p1.ga(1, init_tour=tours[0], epochs=3)              # K8: 3 epochs, population 1000
si = pi.session(T.ALGO_TWO_OPT_BEST, pi.nn_tour(3), T.PATH_MATRIX)  # k1_square (nint) + headline scan
si.enqueue(3)
ctx.sync()
si.close()

n3 = 100000
x3, y3 = bench.gen_uniform(n3, n3)
p3 = T.Problem.euc2d(ctx, x3, y3)
p3.knn(8)                                           # K5 at 100k: 10^10 pairs
print("done")

"""Wall time of tl_knn (kernel + n*k read-back) on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (1000, 10000, 30000):
    x, y = bench.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    for k in (5, 32):
        p.knn(k)
        t0 = time.perf_counter()
        for _ in range(5):
            p.knn(k)
        print(f"knn n={n} k={k}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms wall per call", flush=True)

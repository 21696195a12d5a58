"""Wall time of tl_nn_tour (k-NN lists + walk + tour read-back) on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (1000, 10000, 20000, 50000):
    x, y = bench.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    p.nn_tour(3)
    t0 = time.perf_counter()
    for _ in range(3):
        t = p.nn_tour(3)
    dt = (time.perf_counter() - t0) / 3
    print(f"nn_tour n={n}: {dt * 1e3:.2f} ms wall per call" + (" [TL_NN_NO_HEADS]" if os.environ.get("TL_NN_NO_HEADS") else ""), flush=True)
gx, gy = bench.gen_grid(10000, 10000)
pn = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
pn.nn_tour(3)
t0 = time.perf_counter()
for _ in range(3):
    pn.nn_tour(3)
print(f"nn_tour n=10000 nint (integer grid): {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms wall per call", flush=True)

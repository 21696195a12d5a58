"""n = 50k / 100k, f32 recompute: time for a fixed number of applied moves, full scans vs cached row minima."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
ctx = T.Context(0)
for n, mm in ((50000, 400), (100000, 400)):
    x, y = bench.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    nn = p.nn_tour(3)
    res = {}
    for name, algo in (("full", T.ALGO_TWO_OPT_BEST), ("cached", T.ALGO_TWO_OPT_BEST_CACHED)):
        p.local_search(algo, nn, path=T.PATH_RECOMPUTE, max_moves=3)
        t0 = time.perf_counter()
        t, st, mv = p.local_search(algo, nn, path=T.PATH_RECOMPUTE, max_moves=mm, log_cap=mm)
        wall = time.perf_counter() - t0
        res[name] = (t, mv)
        print(f"n={n} {name}: {mm} moves wall {wall*1e3:.1f} ms device {st.device_ms:.1f} ms evals {int(st.evals):.3e} "
              f"us/step {1e3*st.device_ms/max(1,int(st.passes)):.1f}", flush=True)
    print("  identical moves and tour:", res["full"][1] == res["cached"][1] and bool((res["full"][0] == res["cached"][0]).all()), flush=True)
    p.close()

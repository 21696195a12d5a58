"""Host phases of tl_local_search(OR_OPT) on the 10k nint matrix instance (TL_DEBUG_TIMING=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
ctx = T.Context(0)
x, y = bench.instance(10000, 10000, "nint")
p = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
t2, _, _ = p.local_search(T.ALGO_TWO_OPT_BEST, p.nn_tour(3), path=T.PATH_MATRIX)
for r in range(3):
    t0 = time.perf_counter()
    t3, st, _ = p.local_search(T.ALGO_OR_OPT, t2, path=T.PATH_MATRIX)
    print("or_opt wall ms", 1e3 * (time.perf_counter() - t0), "device ms", st.device_ms, "moves", int(st.moves), "launches", int(st.launches), flush=True)

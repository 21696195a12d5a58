"""Per-step device time of the cached Mode B against the segment length of the move that preceded it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, teeline_b200 as T
torch.cuda.init()
ctx = T.Context(0, stream=torch.cuda.current_stream().cuda_stream)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dist = sys.argv[2] if len(sys.argv) > 2 else "f32"
x, y = bench.instance(n, n, dist)
p = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32 if dist == "nint" else T.DIST_F32_EXACT)
nn = p.nn_tour(3)
s = p.session(T.ALGO_TWO_OPT_BEST_CACHED, nn, T.PATH_MATRIX if dist == "nint" else T.PATH_RECOMPUTE)
K = 400
evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
evs[0].record()
for k in range(K):
    s.enqueue(1)
    evs[k + 1].record()
torch.cuda.synchronize()
ms = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(K)]) * 1e3
log = s.log(K)
L = np.array([m[2] - m[1] for m in log[:K]])
I = np.array([m[1] for m in log[:K]])
print("step 0 (all rows):", ms[0], "us")
# step k+1 re-evaluates what move k changed
prev_L, prev_I, t = L[:K - 1], I[:K - 1], ms[1:K]
for lo, hi in ((0, 8), (8, 64), (64, 512), (512, 2048), (2048, 100000)):
    m = (prev_L >= lo) & (prev_L < hi)
    if m.any():
        print(f"L in [{lo},{hi}): {m.sum()} steps, mean {t[m].mean():.1f} us, median {np.median(t[m]):.1f}, max {t[m].max():.1f}; mean I {prev_I[m].mean():.0f}")
print("total", ms.sum() / 1e3, "ms for", K, "steps")

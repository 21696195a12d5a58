"""Profiling target (run under ncu on the GPU box): `prof_target.py <path> <n> <steps> [algo]`
creates one device-resident session and enqueues <steps> scan+apply steps.  Nothing else runs, so
every kernel in the ncu launch list belongs to the hot path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import teeline_b200 as T  # noqa: E402

path = {"recompute": T.PATH_RECOMPUTE, "matrix": T.PATH_MATRIX}[sys.argv[1]]
n, steps = int(sys.argv[2]), int(sys.argv[3])
algo = {"best": T.ALGO_TWO_OPT_BEST, "oropt": T.ALGO_OR_OPT, "ref": T.ALGO_TWO_OPT_REF}[sys.argv[4] if len(sys.argv) > 4 else "best"]
kind = T.DIST_NINT_I32 if (len(sys.argv) > 5 and sys.argv[5] == "nint") else T.DIST_F32_EXACT
ctx = T.Context(0)
x, y = bench.gen_uniform(n, n) if kind == T.DIST_F32_EXACT else bench.gen_grid(n, n)
p = T.Problem.euc2d(ctx, x, y, kind)
s = p.session(algo, p.nn_tour(3), path)
s.enqueue(steps)
ctx.sync()
st = s.stats()
print({"moves": int(st.moves), "passes": int(st.passes), "device_ms": st.device_ms, "launches": int(st.launches)})

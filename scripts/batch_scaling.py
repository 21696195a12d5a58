"""Config 5 on ONE GPU: time the per-rank share of the 1024 x 1000-city population for world sizes
1, 2, 4, 8 (rank 0's shard; there is no data-path collective, so this IS the per-rank data-path time)
and report which launch configuration the library picked.  usage: batch_scaling.py [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import teeline_b200 as T  # noqa: E402
from teeline_b200 import multi  # noqa: E402

n, B = 1000, 1024
x, y = bench.gen_uniform(n, n)
ctx = T.Context(0)
p = T.Problem.euc2d(ctx, x, y)
tours = np.stack([p.nn_tour(3)] + [bench.shuffle_tour(n, s) for s in range(1, B)])
p.two_opt_batch(tours[:8], max_moves=2)
out = {}
base = None
for world in [int(w) for w in os.environ.get("WORLDS", "1,2,4,8").split(",")]:
    worst, evals = 0.0, 0
    for rank in range(world):
        lo, hi = multi.shard_range(B, rank, world)
        best = 1e9
        for _ in range(2):
            ctx.sync()
            t0 = time.perf_counter()
            _, st, _ = p.two_opt_batch(tours[lo:hi])
            best = min(best, time.perf_counter() - t0)
        worst = max(worst, best)
        evals += int(st.evals)
    base = base or worst
    out[f"world{world}"] = {"tours_per_rank": B // world, "slowest_rank_wall_s": worst, "speedup_vs_1": base / worst,
                            "evals_per_s_aggregate": evals / worst}
    print(world, out[f"world{world}"], flush=True)
if len(sys.argv) > 1:
    os.makedirs(os.path.dirname(os.path.abspath(sys.argv[1])), exist_ok=True)
    json.dump(out, open(sys.argv[1], "w"), indent=1)

"""ncu target, steady-state regime: tours that are 2-opt optimal up to one random segment reversal."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, teeline_b200 as T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 1000
x, y = bench.gen_uniform(n, n)
ctx = T.Context(0)
p = T.Problem.euc2d(ctx, x, y)
tours = bench.shuffle_tours(n, range(1, B + 1))
os.environ["TL_BATCH_ENGINE_SAVE"] = os.environ.get("TL_BATCH_ENGINE", "")
os.environ["TL_BATCH_ENGINE"] = "cta"
opt, _, _ = p.two_opt_batch(tours)  # converge with the CTA-per-tour engine (not profiled: different kernel name)
rng = np.random.default_rng(3)
for b in range(B):
    i, j = sorted(rng.integers(1, n - 1, 2))
    opt[b, i:j + 1] = opt[b, i:j + 1][::-1].copy()
os.environ["TL_BATCH_ENGINE"] = os.environ["TL_BATCH_ENGINE_SAVE"] or "pop"
got, st, lengths = p.two_opt_batch(opt)
print(os.environ["TL_BATCH_ENGINE"], "moves", int(st.moves), "scans", int(st.passes), "device_ms", st.device_ms,
      "evals/s", int(st.evals) / (st.device_ms * 1e-3))

"""Config 5 on one GPU, emulating the per-GPU share of a 1/2/4/8-GPU run (1024/512/256/128 tours of
1000 cities to the local optimum): device time of tl_two_opt_batch per engine / cluster size."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O, teeline_b200 as T
ctx = T.Context(0)
n = 1000
x, y = O.gen_uniform(n, n)
p = T.Problem.euc2d(ctx, x, y)
P = O.Problem(x, y)
tours = np.stack([O.nn_tour(P, 3)] + [O.shuffle_tour(n, s) for s in range(1, 1024)])
variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "auto"]
out = {}
ref = None
for B in (1024, 512, 256, 128, 64, 16):
    for v in variants:
        os.environ.pop("TL_BATCH_CLUSTER", None)
        if v != "auto":
            os.environ["TL_BATCH_CLUSTER"] = v
        p.two_opt_batch(tours[:B], T.ALGO_TWO_OPT_BEST, max_moves=3)
        got, st, lengths = p.two_opt_batch(tours[:B], T.ALGO_TWO_OPT_BEST)
        if B == 1024 and ref is None:
            ref = got.copy()
        assert (got == ref[:B]).all(), (B, v)
        key = f"B{B}_cluster{v}"
        out[key] = {"device_ms": st.device_ms, "moves": int(st.moves), "Tmoves_per_s": int(st.evals) / st.device_ms / 1e9}
        print(key, out[key], flush=True)
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)

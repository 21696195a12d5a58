"""GA wall time, GPU (tl_ga) vs the CPU oracle port, same seed (identical result)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O, teeline_b200 as T
ctx = T.Context(0)
for n, epochs in ((52, 10000), (280, 2000), (1000, 200), (1000, 2000), (4000, 50)):
    if n == 52:
        _, x, y = O.read_tsplib_coords(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "berlin52.tsp"))
    else:
        x, y = O.gen_uniform(n, n)
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    init = O.shuffle_tour(n, 5)
    prob.ga(1, init_tour=init, epochs=2)
    t0 = time.perf_counter(); gt, gc, st = prob.ga(1, init_tour=init, epochs=epochs); tg = time.perf_counter() - t0
    row = {"n": n, "epochs": epochs, "gpu_wall_s": tg, "gpu_device_ms": st.device_ms, "us_per_epoch": 1e3 * st.device_ms / epochs,
           "best": gc, "children_per_s_gpu": int(st.evals) / tg, "launches": int(st.launches)}
    if n * n * epochs <= 1000 * 1000 * 200:
        t0 = time.perf_counter(); ot, oc, _ = O.ga(P, 1, init_tour=init, epochs=epochs); tc = time.perf_counter() - t0
        row.update({"cpu_port_wall_s": tc, "identical": bool((ot == gt.astype(np.int64)).all() and np.float32(oc) == np.float32(gc))})
    print(row, flush=True)

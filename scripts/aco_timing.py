"""ACO wall time, GPU (tl_aco) vs the CPU oracle port, same seed (identical result)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O, teeline_b200 as T
ctx = T.Context(0)
for n, epochs, ants in ((52, 150, 25), (1000, 20, 25), (1000, 20, 148), (2000, 10, 296)):
    if n == 52:
        _, x, y = O.read_tsplib_coords(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "berlin52.tsp"))
    else:
        x, y = O.gen_uniform(n, n)
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    init = O.nn_tour(P, 3)
    prob.aco(1, init_tour=init, epochs=1, num_ants=ants)
    t0 = time.perf_counter(); gt, gc, st = prob.aco(1, init_tour=init, epochs=epochs, num_ants=ants); tg = time.perf_counter() - t0
    row = {"n": n, "epochs": epochs, "ants": ants, "gpu_wall_s": tg, "gpu_device_ms": st.device_ms, "best": gc,
           "weights_per_s_gpu": int(st.evals) / tg, "launches": int(st.launches)}
    if n * ants * epochs <= 1000 * 148 * 20:
        t0 = time.perf_counter(); ot, oc, _ = O.aco(P, 1, init_tour=init, epochs=epochs, num_ants=ants); tc = time.perf_counter() - t0
        row.update({"cpu_port_wall_s": tc, "identical": bool((ot == gt.astype(np.int64)).all() and np.float32(oc) == np.float32(gc))})
    print(row, flush=True)

"""Scan-kernel time of every (algorithm, path, metric) combination at one size (GPU box).
usage: algo_probe.py [n]   -> one JSON line per combination: scan ms, candidates per scan, rate"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import teeline_b200 as T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ctx = T.Context(0)
P2 = (n - 3) * (n - 2) // 2
E_or = n * (n - 2) + 2 * (n - 1) * (n - 3) + 2 * (n - 2) * (n - 4)
for algo_name, algo, cand in (("two_opt_best", T.ALGO_TWO_OPT_BEST, P2), ("or_opt", T.ALGO_OR_OPT, E_or)):
    for path_name, path, kind in (("recompute", T.PATH_RECOMPUTE, T.DIST_F32_EXACT), ("matrix_f32", T.PATH_MATRIX, T.DIST_F32_EXACT),
                                  ("matrix_nint", T.PATH_MATRIX, T.DIST_NINT_I32)):
        x, y = bench.gen_uniform(n, n) if kind == T.DIST_F32_EXACT else bench.gen_grid(n, n)
        p = T.Problem.euc2d(ctx, x, y, kind)
        s = p.session(algo, p.nn_tour(3), path)
        s.enqueue(3)
        ctx.sync()
        import time
        t0 = time.perf_counter()
        s.enqueue(50)
        ctx.sync()
        step_ms = (time.perf_counter() - t0) * 1e3 / 50  # scan + apply (+ rowinfo for Or-opt), host-timed over 50 steps
        ms = s.time_scans(20)
        print(json.dumps({"algo": algo_name, "path": path_name, "n": n, "scan_ms": round(ms, 4), "step_ms": round(step_ms, 4),
                          "candidates": cand, "G_candidates_per_s": round(cand / ms / 1e6, 1)}), flush=True)
        s.close()
        p.close()

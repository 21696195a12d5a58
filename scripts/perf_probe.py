"""Scratch perf probe (run on the GPU box): scan-kernel time and scan+apply step time per build
variant.  usage: perf_probe.py [path:]n ...   e.g.  perf_probe.py 10000 matrix:10000 nint:10000"""
import sys, os, glob, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    import bench
    import teeline_b200 as T
    torch.cuda.init()
    ctx = T.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    out = {}
    for a in sys.argv[2:]:
        kind, n = (a.split(":") + [None])[:2] if ":" in a else ("recompute", a)
        n = int(n)
        if kind == "nint":
            x, y = bench.gen_grid(n, n)
            p = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
        else:
            x, y = bench.gen_uniform(n, n)
            p = T.Problem.euc2d(ctx, x, y)
        path = T.PATH_RECOMPUTE if kind == "recompute" else T.PATH_MATRIX
        s = p.session(T.ALGO_TWO_OPT_BEST, p.nn_tour(3), path)
        pairs = (n - 3) * (n - 2) // 2
        steps = int(os.environ.get('PROBE_STEPS', 100)) if n <= 30000 else 5
        s.enqueue(5); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.enqueue(steps); e1.record(); torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / steps
        scan_ms = s.time_scans(steps)
        out[a] = (round(scan_ms * 1e3, 1), round(pairs / scan_ms / 1e9, 3), round(step_ms * 1e3, 1), round(pairs / step_ms / 1e9, 3))
        s.close()
    print(json.dumps(out))
else:
    libs = [None] + [e for e in os.environ.get("PROBE_ENVS", "").split(",") if e] + sorted(glob.glob(os.path.join(ROOT, "variants", "*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib and lib.endswith(".so"):
            env["TL_LIB"] = lib
        elif lib:  # "NAME" or "NAME=V;NAME2=V2"
            for kv in lib.split(";"):
                k, _, v = kv.partition("=")
                env[k] = v or "1"
        r = subprocess.run([sys.executable, __file__, "child"] + (sys.argv[1:] or ["10000", "100000"]), env=env, capture_output=True, text=True)
        print(os.path.basename(lib) if lib else "default", r.stdout.strip() or r.stderr[-400:], flush=True)
    print("columns per case: scan_us, scan Tmove/s, step_us, step Tmove/s")

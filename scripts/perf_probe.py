"""Scratch perf probe (run on the GPU box): scan-kernel time and scan+apply step time per build variant."""
import sys, os, time, glob, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    import bench
    import teeline_b200 as T
    torch.cuda.init()
    ctx = T.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    out = {}
    for n in [int(a) for a in sys.argv[2:]]:
        x, y = bench.gen_uniform(n, n)
        p = T.Problem.euc2d(ctx, x, y)
        s = p.session(T.ALGO_TWO_OPT_BEST, p.nn_tour(3), T.PATH_RECOMPUTE)
        pairs = (n - 3) * (n - 2) // 2
        steps = 100 if n <= 30000 else 5
        s.enqueue(5); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.enqueue(steps); e1.record(); torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / steps
        scan_ms = s.time_scans(steps)
        out[n] = (round(scan_ms * 1e3, 1), round(pairs / scan_ms / 1e9, 3), round(step_ms * 1e3, 1), round(pairs / step_ms / 1e9, 3))
        s.close()
    print(json.dumps(out))
else:
    libs = [None, "TL_NO_SCREEN"] + sorted(glob.glob(os.path.join(ROOT, "variants", "*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib == "TL_NO_SCREEN":
            env["TL_NO_SCREEN"] = "1"
        elif lib:
            env["TL_LIB"] = lib
        r = subprocess.run([sys.executable, __file__, "child"] + (sys.argv[1:] or ["10000", "100000"]), env=env, capture_output=True, text=True)
        print(os.path.basename(lib) if lib else "default", r.stdout.strip() or r.stderr[-400:], flush=True)
    print("columns per n: scan_us, scan Tmove/s, step_us, step Tmove/s")

"""10k / 100k recompute Mode B: scan-kernel and step time of the library in TL_LIB (tuning variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, teeline_b200 as T
torch.cuda.init()
ctx = T.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for n in (10000, 100000):
    x, y = bench.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    s = p.session(T.ALGO_TWO_OPT_BEST, p.nn_tour(3), T.PATH_RECOMPUTE)
    s.enqueue(10); torch.cuda.synchronize()
    k = 300 if n == 10000 else 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.enqueue(k); e1.record(); torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / k
    scan = s.time_scans(200 if n == 10000 else 20)
    P = bench.pairs_per_scan(n)
    print(os.environ.get("TL_LIB", "default"), "n", n, "step_us %.2f" % (step * 1e3), "scan_us %.2f" % (scan * 1e3), "Tmove/s(step) %.3f" % (P / step / 1e9), flush=True)
    s.close(); p.close()

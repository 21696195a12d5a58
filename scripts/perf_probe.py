"""Scratch perf probe (run on the GPU box): times scan+apply steps and the microbenchmarks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
import teeline_b200 as T

torch.cuda.init()
stream = torch.cuda.current_stream().cuda_stream
ctx = T.Context(0, stream=stream)
ff, mu = ctx.microbench_fp32()
print(f"FFMA lane-instr/s {ff:.3e}  MUFU lane-instr/s {mu:.3e}")
for n in (1000, 10000, 30000, 100000):
    x, y = O.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    t = O.shuffle_tour(n, 1)
    s = p.session(T.ALGO_TWO_OPT_BEST, t, T.PATH_RECOMPUTE)
    pairs = (n - 3) * (n - 2) // 2
    steps = 50 if n <= 30000 else 5
    s.enqueue(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.enqueue(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"n={n}: {ms*1e3:.1f} us/step  {pairs/ms/1e9*1e3/1e3:.3f} Tmove/s (scan+apply)")
    s.close()

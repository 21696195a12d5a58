"""Step timeline of the fused matrix-path step (GPU box; needs a -DTL_TIMELINE build: TL_LIB=variants/lib_tline.so).
usage: TL_LIB=variants/lib_tline.so python scripts/timeline_probe.py [n] [steps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import teeline_b200 as T  # noqa: E402
from teeline_b200 import _capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
torch.cuda.init()
ctx = T.Context(0, stream=torch.cuda.current_stream().cuda_stream)
x, y = bench.gen_grid(n, n)
p = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
s = p.session(T.ALGO_TWO_OPT_BEST, p.nn_tour(3), T.PATH_MATRIX)
lib = _capi.load()
lib.tl_debug_timeline.argtypes = [C.POINTER(C.c_double), C.c_int]
out = (C.c_double * 8)()
s.enqueue(10)
torch.cuda.synchronize()
lib.tl_debug_timeline(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.enqueue(steps); e1.record(); torch.cuda.synchronize()
lib.tl_debug_timeline(out, 0)
cnt = out[7]
names = ["gap prev done -> first CTA past wait", "first CTA scan end", "last CTA scan end", "tail entry (last CTA)",
         "candidates reduced", "segment reversed", "step done"]
print(f"n={n} steps={steps} counted={int(cnt)}  event-timed step {e0.elapsed_time(e1) / steps * 1e3:.2f} us")
for k, nm in enumerate(names):
    print(f"  {nm:40s} {out[k] / cnt / 1e3:8.2f} us" + ("  (since first CTA past wait)" if k else ""))

import numpy as np
raw = (C.c_double * (3 * 1024))()
lib.tl_debug_timeline(raw, 2)
a = np.frombuffer(raw, dtype=np.uint64).reshape(3, 1024)
grid = int(np.count_nonzero(a[1]))
t0 = a[0, :grid].min()
start, end, sm = (a[0, :grid] - t0).astype(np.int64), (a[1, :grid] - t0).astype(np.int64), a[2, :grid].astype(np.int64)
print(f"last step, {grid} CTAs: start past wait min/mean/max {start.min()}/{start.mean():.0f}/{start.max()} ns; "
      f"scan end min/mean/max {end.min()}/{end.mean():.0f}/{end.max()} ns")
order = np.argsort(end)
print("slowest 24 CTAs (block, sm, start, end):", [(int(b), int(sm[b]), int(start[b]), int(end[b])) for b in order[-24:]])
print("fastest 24 CTAs (block, sm, start, end):", [(int(b), int(sm[b]), int(start[b]), int(end[b])) for b in order[:24]])
per_sm = {}
for b in range(grid):
    per_sm.setdefault(int(sm[b]), []).append(int(end[b]))
sm_end = sorted((max(v), k, len(v)) for k, v in per_sm.items())
print("per-SM latest end, slowest 16 (end, sm, ctas):", sm_end[-16:])
print("per-SM latest end, fastest 16:", sm_end[:16])
hist, edges = np.histogram(end, bins=12)
print("end-time histogram:", list(zip(edges[:-1].astype(int).tolist(), hist.tolist())))
print("corr(end, block) =", float(np.corrcoef(end, np.arange(grid))[0, 1]), " corr(end, sm) =", float(np.corrcoef(end, sm)[0, 1]))

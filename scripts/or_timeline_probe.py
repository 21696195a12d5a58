"""Per-warp timeline of one Or-opt scan (GPU box; needs a -DTL_TIMELINE build via TL_LIB)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench, teeline_b200 as T
from teeline_b200 import _capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ctx = T.Context(0)
x, y = bench.gen_uniform(n, n)
p = T.Problem.euc2d(ctx, x, y)
s = p.session(T.ALGO_OR_OPT, p.nn_tour(3), T.PATH_RECOMPUTE)
s.enqueue(3); ctx.sync()
ms = s.time_scans(10)
lib = _capi.load()
buf = (C.c_ulonglong * (3 * 4096))()
lib.tl_debug_or_timeline.argtypes = [C.POINTER(C.c_ulonglong)]
lib.tl_debug_or_timeline(buf)
a = np.frombuffer(buf, dtype=np.uint64).reshape(3, 4096)
w = int(np.count_nonzero(a[1]))
t0 = a[0, :w].min()
st, en, it = (a[0, :w] - t0).astype(np.int64), (a[1, :w] - t0).astype(np.int64), a[2, :w].astype(np.int64)
print(f"scan {ms*1e3:.1f} us (rowinfo + scan launches); {w} warps; start min/mean/max {st.min()}/{st.mean():.0f}/{st.max()} ns; "
      f"end min/mean/max {en.min()}/{en.mean():.0f}/{en.max()} ns; items per warp min/mean/max {it.min()}/{it.mean():.2f}/{it.max()}")
h, e = np.histogram(en, bins=10)
print("end histogram:", list(zip(e[:-1].astype(int).tolist(), h.tolist())))
h, e = np.histogram(st, bins=10)
print("start histogram:", list(zip(e[:-1].astype(int).tolist(), h.tolist())))
dur = en - st
per_item = dur / np.maximum(it, 1)
print("per-item time (us) percentiles 1/10/50/90/99:", [round(float(np.percentile(per_item, q)) / 1e3, 1) for q in (1, 10, 50, 90, 99)])
for k in sorted(set(it.tolist())):
    sel = it == k
    print(f"  warps with {k} items: {int(sel.sum())}, end mean {en[sel].mean() / 1e3:.1f} us, per-item mean {per_item[sel].mean() / 1e3:.1f} us")
cta_end = en.reshape(-1, 8).max(axis=1) if w % 8 == 0 else None
if cta_end is not None:
    print("per-CTA latest end (us) percentiles 1/50/99:", [round(float(np.percentile(cta_end, q)) / 1e3, 1) for q in (1, 50, 99)])
    # do the warps of one CTA finish together?
    spread = en.reshape(-1, 8).max(axis=1) - en.reshape(-1, 8).min(axis=1)
    print("within-CTA end spread (us) mean/max:", round(float(spread.mean()) / 1e3, 1), round(float(spread.max()) / 1e3, 1))

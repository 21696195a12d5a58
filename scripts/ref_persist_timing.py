"""Mode R (reference-exact first improvement) to the local optimum: per-step kernel vs the persistent
cluster kernel (TL_REF_CLUSTER=0 | 8 | 16).  usage: ref_persist_timing.py <tag>  (GPU box)"""
import os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, teeline_b200 as T
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
ctx = T.Context(0)
import ctypes as C
from teeline_b200 import _capi
_lib = _capi.load()
_prof = getattr(_lib, "tl_debug_refp", None)  # tuning builds (-DTL_REFP_PROF) only
if _prof is not None:
    _prof.argtypes = [C.POINTER(C.c_double), C.c_int]
for n, dist in ((52, "f32"), (1000, "f32"), (5000, "f32"), (10000, "f32"), (14000, "f32"), (1000, "nint"), (10000, "nint"), (11000, "nint")):
    x, y = bench.instance(n, n, dist)
    p = T.Problem.euc2d(ctx, x, y, T.DIST_F32_EXACT if dist == "f32" else T.DIST_NINT_I32)
    nn = p.nn_tour(3)
    p.local_search(T.ALGO_TWO_OPT_REF, nn, max_moves=5)
    walls = []
    if _prof is not None:
        o = (C.c_double * 8)()
        _prof(o, 1)
        p.local_search(T.ALGO_TWO_OPT_REF, nn)
        _prof(o, 1)
        steps, hits, units, ev, bar, ap, tot, rows = list(o)
        print(f"[{tag}] n={n} profile (rank 0, thread 0): {int(steps)} steps, {int(hits)} hits, {int(rows)} window rows, "
              f"{int(units)} units by this warp; cycles: eval {ev/1e6:.2f} M, barrier wait {bar/1e6:.2f} M, apply {ap/1e6:.2f} M, "
              f"kernel {tot/1e6:.2f} M; per step: eval {ev/steps:.0f}, barrier {bar/steps:.0f}, apply {ap/steps:.0f}", flush=True)
    for _ in range(3):
        t0 = time.perf_counter()
        t, st, mv = p.local_search(T.ALGO_TWO_OPT_REF, nn, log_cap=1 << 16)
        walls.append(time.perf_counter() - t0)
    h = hashlib.sha1(np.asarray(t, dtype=np.int64).tobytes()).hexdigest()[:12]
    hm = hashlib.sha1(repr([(m[1], m[2]) for m in mv]).encode()).hexdigest()[:12]
    print(f"[{tag}] n={n} {dist}: wall {1e3*min(walls):.2f} ms device {st.device_ms:.2f} ms moves {int(st.moves)} passes {int(st.passes)} "
          f"launches {int(st.launches)} tour {h} log {hm}", flush=True)
    # budgeted + resumed run must give the same tour (cursor state survives a launch boundary)
    s = p.session(T.ALGO_TWO_OPT_REF, nn)
    s.run(max_moves=7)
    s.run(max_moves=int(st.moves) // 2)
    s.run()
    t2 = s.tour()
    st2 = s.stats()
    s.close()
    print(f"[{tag}]   resumed in 3 runs: same tour {bool((np.asarray(t2) == np.asarray(t)).all())} moves {int(st2.moves)} passes {int(st2.passes)}", flush=True)
    p.close()

"""ctypes loader for the CPU ORACLE (oracle/libteeline_oracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package
(teeline_b200/) never imports this module.  See oracle/teeline_oracle.h for the
reference file:line each function restates and for the parity-pinned status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libteeline_oracle.so")

EUC_F32, PACKED_F32, PACKED_I32 = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_DIR, f) for f in ("teeline_oracle.c", "algos.inc", "teeline_oracle.h")]
    stale = not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.run(["make", "-C", _DIR, "-s"], check=True)
    return _SO


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("kind", C.c_int32), ("x", C.c_void_p), ("y", C.c_void_p),
                ("tri_f", C.c_void_p), ("tri_i", C.c_void_p)]


class Move(C.Structure):
    _fields_ = [("delta", C.c_double), ("i", C.c_int32), ("j", C.c_int32),
                ("seg_len", C.c_int32), ("reversed", C.c_int32)]

    def astuple(self):
        return (self.delta, self.i, self.j, self.seg_len, self.reversed)


class Stats(C.Structure):
    _fields_ = [("passes", C.c_int64), ("moves", C.c_int64), ("evals", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.tlo_dist_f32.restype = C.c_float
        L.tlo_dist_f32.argtypes = [C.c_float] * 4
        L.tlo_dist_nint.restype = C.c_int32
        L.tlo_dist_nint.argtypes = [C.c_float] * 4
        L.tlo_distance.restype = C.c_double
        L.tlo_tour_length.restype = C.c_double
        L.tlo_splitmix64.restype = C.c_uint64
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Problem:
    """A problem instance for the oracle: coordinates (EUC_F32) or a packed triangle."""

    def __init__(self, x=None, y=None, tri=None, n=None):
        self._keep = []
        st = _Problem()
        if tri is not None:
            tri = np.ascontiguousarray(tri)
            assert n is not None and tri.size == n * (n - 1) // 2
            if tri.dtype == np.int32:
                st.kind, st.tri_i = PACKED_I32, _p(tri)
            else:
                tri = np.ascontiguousarray(tri, dtype=np.float32)
                st.kind, st.tri_f = PACKED_F32, _p(tri)
            self._keep.append(tri)
            st.n = n
        else:
            x = np.ascontiguousarray(x, dtype=np.float32)
            y = np.ascontiguousarray(y, dtype=np.float32)
            st.kind, st.n, st.x, st.y = EUC_F32, len(x), _p(x), _p(y)
            self._keep += [x, y]
        self.x, self.y = x, y
        self.n = int(st.n)
        self.st = st

    @property
    def ref(self):
        return C.byref(self.st)


def dist_f32(x1, y1, x2, y2) -> np.float32:
    return np.float32(lib().tlo_dist_f32(x1, y1, x2, y2))


def dist_nint(x1, y1, x2, y2) -> int:
    return int(lib().tlo_dist_nint(x1, y1, x2, y2))


def matrix_packed_f32(x, y) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    n = len(x)
    out = np.empty(n * (n - 1) // 2, dtype=np.float32)
    lib().tlo_matrix_packed_f32(C.c_int32(n), _p(x), _p(y), _p(out))
    return out


def matrix_packed_nint(x, y) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    n = len(x)
    out = np.empty(n * (n - 1) // 2, dtype=np.int32)
    lib().tlo_matrix_packed_nint(C.c_int32(n), _p(x), _p(y), _p(out))
    return out


def distance(p: Problem, a: int, b: int) -> float:
    return float(lib().tlo_distance(p.ref, C.c_int32(a), C.c_int32(b)))


def _tour(t):
    return np.ascontiguousarray(t, dtype=np.int32)


def tour_length(p: Problem, tour) -> float:
    t = _tour(tour)
    return float(lib().tlo_tour_length(p.ref, _p(t), C.c_int32(len(t))))


def tour_lengths(p: Problem, tours) -> np.ndarray:
    t = np.ascontiguousarray(tours, dtype=np.int32)
    b, n = t.shape
    out = np.empty(b, dtype=np.float64)
    lib().tlo_tour_lengths(p.ref, _p(t), C.c_int64(b), C.c_int32(n), _p(out))
    return out


def knn(p: Problem, k: int) -> np.ndarray:
    out = np.empty((p.n, k), dtype=np.int32)
    lib().tlo_knn(p.ref, C.c_int32(k), _p(out))
    return out


def nn_tour(p: Problem, k: int = 3) -> np.ndarray:
    out = np.empty(p.n, dtype=np.int32)
    lib().tlo_nn_tour(p.ref, C.c_int32(k), _p(out))
    return out


def swap_2opt(path, frm: int, to: int) -> np.ndarray:
    t = _tour(path).copy()
    lib().tlo_swap_2opt(_p(t), C.c_int32(frm), C.c_int32(to))
    return t


def _log(cap):
    return (Move * max(cap, 1))()


def _moves(log, st, cap):
    return [log[k].astuple() for k in range(min(st.moves, cap))]


def two_opt_ref(p: Problem, tour, log_cap: int = 0):
    """Mode R (reference first-improvement).  Returns (tour, stats, moves)."""
    t = _tour(tour).copy()
    st, log = Stats(), _log(log_cap)
    lib().tlo_two_opt_ref(p.ref, _p(t), C.byref(st), log, C.c_int64(log_cap))
    return t, st, _moves(log, st, log_cap)


def two_opt_best_scan(p: Problem, tour, cyclic: bool = False, nthreads: int = 1):
    t = _tour(tour)
    mv = Move()
    found = lib().tlo_two_opt_best_scan(p.ref, _p(t), C.c_int(int(cyclic)), C.c_int(nthreads),
                                        C.byref(mv))
    return mv.astuple() if found else None


def two_opt_best(p: Problem, tour, cyclic: bool = False, max_moves: int = -1, nthreads: int = 1,
                 log_cap: int = 0):
    """Mode B (best-improvement).  Returns (tour, stats, moves)."""
    t = _tour(tour).copy()
    st, log = Stats(), _log(log_cap)
    lib().tlo_two_opt_best(p.ref, _p(t), C.c_int(int(cyclic)), C.c_int64(max_moves),
                           C.c_int(nthreads), C.byref(st), log, C.c_int64(log_cap))
    return t, st, _moves(log, st, log_cap)


def or_opt_find_best(p: Problem, tour, nthreads: int = 1):
    t = _tour(tour)
    mv = Move()
    if nthreads > 1:
        found = lib().tlo_or_opt_find_best_mt(p.ref, _p(t), C.c_int(nthreads), C.byref(mv), None)
    else:
        found = lib().tlo_or_opt_find_best(p.ref, _p(t), C.byref(mv))
    return mv.astuple() if found else None


def or_opt_apply(tour, i: int, seg_len: int, j: int, reversed_: bool) -> np.ndarray:
    t = _tour(tour).copy()
    lib().tlo_or_opt_apply(_p(t), C.c_int32(len(t)), C.c_int32(i), C.c_int32(seg_len),
                           C.c_int32(j), C.c_int32(int(reversed_)))
    return t


def or_opt(p: Problem, tour, max_moves: int = -1, log_cap: int = 0):
    t = _tour(tour).copy()
    st, log = Stats(), _log(log_cap)
    lib().tlo_or_opt(p.ref, _p(t), C.c_int64(max_moves), C.byref(st), log, C.c_int64(log_cap))
    return t, st, _moves(log, st, log_cap)


def three_opt_find_best(p: Problem, tour, nthreads: int = 1):
    """three_opt.rs:58-131.  Returns (delta = -savings, i, j, k, case) or None, and the triples evaluated."""
    t = _tour(tour)
    mv, k, kase, ev = Move(), C.c_int32(), C.c_int32(), C.c_int64()
    found = lib().tlo_three_opt_find_best(p.ref, _p(t), C.c_int(nthreads), C.byref(mv), C.byref(k),
                                          C.byref(kase), C.byref(ev))
    return ((mv.delta, mv.i, mv.j, k.value, kase.value) if found else None), ev.value


def three_opt_apply(tour, i: int, j: int, k: int, case: int) -> np.ndarray:
    t = _tour(tour).copy()
    lib().tlo_three_opt_apply(_p(t), C.c_int32(i), C.c_int32(j), C.c_int32(k), C.c_int32(case))
    return t


def three_opt(p: Problem, tour, max_moves: int = -1, nthreads: int = 1, log_cap: int = 0):
    """three_opt::solve.  Returns (tour, stats, moves) with moves = (delta, i, j, k, case)."""
    t = _tour(tour).copy()
    st, log = Stats(), _log(log_cap)
    ks = np.zeros(max(log_cap, 1), dtype=np.int32)
    lib().tlo_three_opt(p.ref, _p(t), C.c_int64(max_moves), C.c_int(nthreads), C.byref(st), log, _p(ks),
                        C.c_int64(log_cap))
    mv = [(log[m].delta, log[m].i, log[m].j, int(ks[m]), log[m].seg_len) for m in range(min(st.moves, log_cap))]
    return t, st, mv


class AcoOptions(C.Structure):
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("evaporation_rate", C.c_float),
                ("num_ants", C.c_int32), ("epochs", C.c_int32), ("seed", C.c_uint64)]


def philox4x32(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().tlo_philox4x32(c, k, out)
    return list(out)


def aco(p: Problem, seed: int, init_tour=None, alpha=1.0, beta=2.0, evaporation_rate=0.5, num_ants=25, epochs=150):
    """Ant System with the reference's defaults (mod.rs:1091-1111).  Returns (best_tour, best_cost, stats)."""
    o = AcoOptions(alpha, beta, evaporation_rate, num_ants, epochs, seed)
    best = np.empty(p.n, dtype=np.int32)
    st = Stats()
    it = _tour(init_tour) if init_tour is not None else None
    lib().tlo_aco.restype = C.c_double
    cost = lib().tlo_aco(p.ref, C.byref(o), _p(it) if it is not None else None, _p(best), C.byref(st))
    return best, float(cost), st


class GaOptions(C.Structure):
    _fields_ = [("mutation_probability", C.c_float), ("n_elite", C.c_int32), ("epochs", C.c_int32),
                ("pad", C.c_int32), ("seed", C.c_uint64)]


def ox_genes(p1, p2, frm: int, to: int):
    """ordered_crossover_genes (genetic_algorithm.rs:140-176)."""
    a, b = _tour(p1), _tour(p2)
    g1, g2 = np.empty_like(a), np.empty_like(b)
    lib().tlo_ox_genes(_p(a), _p(b), C.c_int32(len(a)), C.c_int32(frm), C.c_int32(to), _p(g1), _p(g2))
    return g1, g2


def ga(p: Problem, seed: int, init_tour=None, mutation_probability=0.001, n_elite=3, epochs=10000):
    """GA with the reference's GAOptions defaults.  Returns (best_tour, best_length, stats)."""
    o = GaOptions(mutation_probability, n_elite, epochs, 0, seed)
    best = np.empty(p.n, dtype=np.int32)
    st = Stats()
    it = _tour(init_tour) if init_tour is not None else None
    lib().tlo_ga.restype = C.c_double
    cost = lib().tlo_ga(p.ref, C.byref(o), _p(it) if it is not None else None, _p(best), C.byref(st))
    return best, float(cost), st


def gen_uniform(n: int, seed: int):
    x = np.empty(n, dtype=np.float32)
    y = np.empty(n, dtype=np.float32)
    lib().tlo_gen_uniform(C.c_int32(n), C.c_uint64(seed), _p(x), _p(y))
    return x, y


def gen_grid(n: int, seed: int):
    x = np.empty(n, dtype=np.float32)
    y = np.empty(n, dtype=np.float32)
    lib().tlo_gen_grid(C.c_int32(n), C.c_uint64(seed), _p(x), _p(y))
    return x, y


def shuffle_tour(n: int, seed: int) -> np.ndarray:
    t = np.empty(n, dtype=np.int32)
    lib().tlo_shuffle_tour(C.c_int32(n), C.c_uint64(seed), _p(t))
    return t


def read_tsplib_coords(path: str, cap: int = 1 << 20):
    ids = np.empty(cap, dtype=np.int64)
    x = np.empty(cap, dtype=np.float32)
    y = np.empty(cap, dtype=np.float32)
    n = lib().tlo_read_tsplib_coords(path.encode(), C.c_int32(cap), _p(ids), _p(x), _p(y))
    if n < 0:
        raise IOError(f"tlo_read_tsplib_coords({path}) -> {n}")
    return ids[:n].copy(), x[:n].copy(), y[:n].copy()


def read_opt_tour(path: str) -> np.ndarray:
    """TSPLIB .opt.tour reader (src/tsp/opt_tour.rs): ids after TOUR_SECTION until -1/EOF."""
    out, on = [], False
    with open(path) as f:
        for line in f:
            s = line.strip().upper()
            if s == "TOUR_SECTION":
                on = True
                continue
            if not on or not s:
                continue
            if s == "EOF" or s == "-1":
                break
            out += [int(tok) for tok in s.split() if tok != "-1"]
    return np.array(out, dtype=np.int64)

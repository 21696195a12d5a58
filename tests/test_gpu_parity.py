"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
reference's published goldens.  Bit-exact for everything: distances, move sequences, tours."""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import teeline_b200 as T
    return T


@pytest.fixture(scope="module")
def ctx(T):
    c = T.Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def f5(v):
    return "%.5f" % v


# ---- numerics foundation ---------------------------------------------------------------

def test_fast_sqrt_is_ieee_over_its_whole_domain(ctx):
    """sqrt_rn_fast == sqrt.rn for EVERY f32 in [2^-101, FLT_MAX] and for +0 (exhaustive)."""
    assert ctx.selftest_sqrt(0x0D000000, 0x7F7FFFFF) == 0
    assert ctx.selftest_sqrt(0, 0) == 0


# ---- K1 -------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["berlin52.tsp", "a280.tsp", "att532.tsp"])
def test_k1_packed_matches_oracle_fixtures(T, ctx, golden_dir, name):
    _, x, y = O.read_tsplib_coords(os.path.join(golden_dir, name))
    p = T.Problem.euc2d(ctx, x, y)
    assert (bits(p.matrix_packed()) == bits(O.matrix_packed_f32(x, y))).all()


@pytest.mark.parametrize("n", [2, 3, 5, 1000, 3001])
def test_k1_packed_matches_oracle_synthetic(T, ctx, n):
    x, y = O.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    assert (bits(p.matrix_packed()) == bits(O.matrix_packed_f32(x, y))).all()


def test_k1_reference_unit_vectors(T, ctx):  # distance_matrix.rs:326-349
    p = T.Problem.euc2d(ctx, [0.0, 0.0], [0.0, 1.0])
    assert p.matrix_packed().tolist() == [1.0]
    m = T.Problem.euc2d(ctx, [0.0, 0.0, 2.0], [0.0, 1.0, 0.0]).matrix_packed()
    assert m[0] == 1.0 and m[1] == 2.0 and abs(m[2] - 2.236068) < 1e-6
    with pytest.raises(T.TeelineError):  # "distance matrix requires at least 2 points"
        T.Problem.euc2d(ctx, [100.0], [100.0])


def test_k1_nint_matches_oracle(T, ctx, berlin52):
    _, x, y = berlin52
    p = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
    assert (p.matrix_packed() == O.matrix_packed_nint(x, y)).all()
    gx, gy = O.gen_grid(1500, 1500)
    p = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
    assert (p.matrix_packed() == O.matrix_packed_nint(gx, gy)).all()


@pytest.mark.parametrize("case", ["wide_ints", "offset_ints", "half_integers", "long_thin"])
def test_k1_nint_coordinate_ranges(T, ctx, case):
    """TSPLIB nint over coordinate sets that stress the rounding: distances up to 1.5e6, large
    offsets, half-integers, d2 = 0, 1, 4, 9, ... neighbours."""
    rng = np.random.default_rng(7)
    n = 1200
    if case == "wide_ints":              # ranges 2^20 wide, distances up to 1.48e6
        x = rng.integers(0, 1048577, n).astype(np.float32)
        y = rng.integers(0, 1048577, n).astype(np.float32)
        x[:2], y[:2] = [0, 1048576], [0, 1048576]
    elif case == "offset_ints":          # integer coordinates near 2^22
        x = (3145728 + rng.integers(0, 1000000, n)).astype(np.float32)
        y = (-4194304 + rng.integers(0, 1000000, n)).astype(np.float32)
    elif case == "half_integers":
        x = (rng.integers(0, 2000, n) + 0.5).astype(np.float32)
        y = rng.integers(0, 2000, n).astype(np.float32)
    else:
        x = rng.integers(0, 4000000, n).astype(np.float32)
        y = rng.integers(0, 1000, n).astype(np.float32)
    # neighbours at small distances too
    x[10:20], y[10:20] = x[0] + np.arange(10, dtype=np.float32), y[0]
    x[20], y[20] = x[0], y[0]
    p = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
    assert (p.matrix_packed() == O.matrix_packed_nint(x, y)).all()


def test_k1_safe_sqrt_path_for_wild_coordinates(T, ctx):
    """Coordinates outside the fast-sqrt guarantee (tiny / huge) take the IEEE-safe kernels."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(300) * 1e-30).astype(np.float32)
    y = (rng.standard_normal(300) * 1e-30).astype(np.float32)
    x[:10] = 0.0
    p = T.Problem.euc2d(ctx, x, y)
    assert (bits(p.matrix_packed()) == bits(O.matrix_packed_f32(x, y))).all()


# ---- K4 -------------------------------------------------------------------------------

def test_k4_goldens(T, ctx, berlin52, golden_dir):
    ids, x, y = berlin52
    p = T.Problem.euc2d(ctx, x, y)
    P = O.Problem(x, y)
    nn = O.nn_tour(P, 3)
    assert f5(p.tour_lengths(nn)[0]) == "8980.91797"  # G1
    opt = O.read_opt_tour(os.path.join(golden_dir, "berlin52.opt.tour"))
    pos = {int(c): k for k, c in enumerate(ids)}
    assert f5(p.tour_lengths([pos[int(c)] for c in opt])[0]) == "7544.36572"  # G3


def test_k4_batch_exact_matches_oracle(T, ctx):
    n, B = 1000, 257
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    p = T.Problem.euc2d(ctx, x, y)
    tours = np.stack([O.shuffle_tour(n, s) for s in range(1, B + 1)])
    want = O.tour_lengths(P, tours).astype(np.float32)
    got = p.tour_lengths(tours, T.LEN_EXACT)
    assert (bits(got) == bits(want)).all()
    fast = p.tour_lengths(tours, T.LEN_FAST)
    assert np.allclose(fast, want, rtol=1e-5)  # f64 pairwise sum of the same f32 edges


def test_k4_edge_cases(T, ctx):
    x, y = O.gen_uniform(33, 3)
    P = O.Problem(x, y)
    p = T.Problem.euc2d(ctx, x, y)
    t = O.shuffle_tour(33, 9)
    assert bits(p.tour_lengths(t))[0] == bits(np.float32(O.tour_length(P, t)))[0]
    bad = t.copy().astype(np.uint32)
    bad[7] = 1000  # unknown position -> 0.0 (distance_matrix.rs:221-231)
    assert p.tour_lengths(bad)[0] == 0.0
    p2 = T.Problem.euc2d(ctx, [0.0, 3.0], [0.0, 4.0])
    assert p2.tour_lengths([0, 1])[0] == 10.0  # closing edge + the one window


def test_k4_explicit_and_nint(T, ctx, berlin52, golden_dir):
    ids, x, y = berlin52
    tri = O.matrix_packed_f32(x, y)
    pe = T.Problem.explicit(ctx, tri, 52)
    P = O.Problem(x, y)
    nn = O.nn_tour(P, 3)
    assert f5(pe.tour_lengths(nn)[0]) == "8980.91797"
    pi = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
    opt = O.read_opt_tour(os.path.join(golden_dir, "berlin52.opt.tour"))
    pos = {int(c): k for k, c in enumerate(ids)}
    assert pi.tour_lengths([pos[int(c)] for c in opt])[0] == 7542


# ---- K2 Mode B, recompute path ------------------------------------------------------------

def check_mode_b(T, ctx, x, y, start, cyclic=False, max_moves=-1):
    P = O.Problem(x, y)
    want_t, want_st, want_mv = O.two_opt_best(P, start, cyclic=cyclic, max_moves=max_moves, nthreads=4,
                                              log_cap=1 << 16)
    p = T.Problem.euc2d(ctx, x, y)
    algo = T.ALGO_TWO_OPT_BEST_CYCLIC if cyclic else T.ALGO_TWO_OPT_BEST
    got_t, st, mv = p.local_search(algo, start, path=T.PATH_RECOMPUTE, max_moves=max_moves, log_cap=1 << 16)
    assert [(m[1], m[2]) for m in mv] == [(m[1], m[2]) for m in want_mv]
    assert [np.float32(m[0]) for m in mv] == [np.float32(m[0]) for m in want_mv]
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)
    assert bool(st.converged) == (max_moves < 0 or want_st.moves < max_moves)
    return got_t, st


def test_k2_best_berlin52(T, ctx, berlin52):
    _, x, y = berlin52
    check_mode_b(T, ctx, x, y, O.nn_tour(O.Problem(x, y), 3))
    check_mode_b(T, ctx, x, y, np.arange(52))
    check_mode_b(T, ctx, x, y, np.arange(52), cyclic=True)


@pytest.mark.parametrize("n", [4, 5, 6, 7, 13, 290, 291, 292, 600])
def test_k2_best_small_and_band_edges(T, ctx, n):
    x, y = O.gen_uniform(n, 100 + n)
    for cyclic in (False, True):
        check_mode_b(T, ctx, x, y, O.shuffle_tour(n, n), cyclic=cyclic)


def test_k2_best_tiny_n_is_noop(T, ctx):
    for n in (2, 3):
        x, y = O.gen_uniform(n, n)
        p = T.Problem.euc2d(ctx, x, y)
        t, st, mv = p.local_search(T.ALGO_TWO_OPT_BEST, np.arange(n)[::-1].copy())
        assert t.tolist() == list(range(n))[::-1] and int(st.moves) == 0 and st.converged == 1


def test_k2_best_1k_nn_start_matches_survey_probe(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    P = O.Problem(x, y)
    t, st = check_mode_b(T, ctx, x, y, O.nn_tour(P, 3))
    assert f5(O.tour_length(P, t)) == "25282.04297" and int(st.moves) == 170


def test_k2_best_1k_random_start_bounded(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    check_mode_b(T, ctx, x, y, O.shuffle_tour(1000, 11), max_moves=60)


def test_k2_best_ties_resolve_to_lowest_ij(T, ctx):
    """Integer lattice => many exactly equal deltas; the lowest (i,j) must win every time."""
    rng = np.random.default_rng(3)
    x = rng.integers(0, 12, 400).astype(np.float32)
    y = rng.integers(0, 12, 400).astype(np.float32)
    check_mode_b(T, ctx, x, y, O.shuffle_tour(400, 2), max_moves=80)
    check_mode_b(T, ctx, x, y, O.shuffle_tour(400, 2), cyclic=True, max_moves=80)


def test_k2_scan_only_10k(T, ctx):
    """One full-size scan (P(10k) = 49 975 003 pairs) against the threaded oracle scan."""
    n = 10000
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    p = T.Problem.euc2d(ctx, x, y)
    for seed in (1, 2):
        t = O.shuffle_tour(n, seed)
        s = p.session(T.ALGO_TWO_OPT_BEST, t, T.PATH_RECOMPUTE)
        got = s.scan()
        want = O.two_opt_best_scan(P, t, nthreads=8)
        assert got is not None and (got[1], got[2]) == (want[1], want[2])
        assert np.float32(got[0]) == np.float32(want[0])
        s.close()


def test_k2_rejects_bad_tours(T, ctx):
    x, y = O.gen_uniform(10, 1)
    p = T.Problem.euc2d(ctx, x, y)
    with pytest.raises(T.TeelineError):
        p.local_search(T.ALGO_TWO_OPT_BEST, [0, 1, 2, 3, 4, 5, 6, 7, 8, 8])
    with pytest.raises(T.TeelineError):
        p.local_search(T.ALGO_TWO_OPT_BEST, [0, 1, 2, 3, 4, 5, 6, 7, 8, 10])


# ---- K2 Mode R (reference-exact first improvement) -------------------------------------------

def check_mode_r(T, ctx, x, y, start, max_moves=-1):
    P = O.Problem(x, y)
    want_t, want_st, want_mv = O.two_opt_ref(P, start, log_cap=1 << 16)
    p = T.Problem.euc2d(ctx, x, y)
    got_t, st, mv = p.local_search(T.ALGO_TWO_OPT_REF, start, log_cap=1 << 16)
    assert [(m[1], m[2]) for m in mv] == [(m[1], m[2]) for m in want_mv]
    assert [np.float32(m[0]) for m in mv] == [np.float32(m[0]) for m in want_mv]
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)
    assert st.converged == 1
    return got_t, st


def test_mode_r_goldens_G4_G5(T, ctx, berlin52):
    _, x, y = berlin52
    P = O.Problem(x, y)
    p = T.Problem.euc2d(ctx, x, y)
    t, st = check_mode_r(T, ctx, x, y, np.arange(52))
    assert f5(p.tour_lengths(t)[0]) == "9368.31836"  # G4, docs/benchmarks.md:28
    assert (int(st.moves), int(st.passes), int(st.evals)) == (87, 5, 6125)
    t, st = check_mode_r(T, ctx, x, y, O.nn_tour(P, 3))
    assert f5(p.tour_lengths(t)[0]) == "8384.18848"  # G5, README.md:385
    assert (int(st.moves), int(st.passes), int(st.evals)) == (8, 3, 3675)


def test_mode_r_tsp5_reference_unit_test(T, ctx):  # two_opt.rs:100-131
    x, y = [0.0, 0.0, 0.0, 1.0, 1.0], [0.0, 0.5, 1.0, 1.0, 0.0]
    p = T.Problem.euc2d(ctx, x, y)
    t, st, _ = p.local_search(T.ALGO_TWO_OPT_REF, np.arange(5))
    assert t.tolist() == [0, 1, 2, 3, 4] and p.tour_lengths(t)[0] == 4.0


@pytest.mark.parametrize("n", [3, 4, 5, 6, 9, 40, 300])
def test_mode_r_small(T, ctx, n):
    x, y = O.gen_uniform(n, 200 + n)
    check_mode_r(T, ctx, x, y, O.shuffle_tour(n, n + 1))


def test_mode_r_1k_matches_survey_probe(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    P = O.Problem(x, y)
    t, st = check_mode_r(T, ctx, x, y, O.nn_tour(P, 3))
    assert f5(O.tour_length(P, t)) == "25436.69336"
    assert (int(st.passes), int(st.moves), int(st.evals)) == (6, 331, 2985018)


def test_mode_r_1k_random_start(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    check_mode_r(T, ctx, x, y, O.shuffle_tour(1000, 5))


def test_mode_r_move_budgets_resume_where_they_stopped(T, ctx):
    """The persistent cluster kernel keeps cursor, window and pass bookkeeping in registers and hands
    them back through DevState: a search cut into budgeted runs applies the oracle's moves in the
    oracle's order and ends in the oracle's tour."""
    n = 2000
    x, y = O.gen_uniform(n, 77)
    P = O.Problem(x, y)
    start = O.nn_tour(P, 3)
    want_t, want_st, want_mv = O.two_opt_ref(P, start, log_cap=1 << 16)
    p = T.Problem.euc2d(ctx, x, y)
    s = p.session(T.ALGO_TWO_OPT_REF, start)
    prefix = np.asarray(start, dtype=np.int64).copy()
    done = 0
    for budget in (1, 7, 8, want_st.moves // 2, want_st.moves - 1):
        s.run(max_moves=budget)
        for m in want_mv[done:budget]:
            prefix[m[1] + 1:m[2] + 1] = prefix[m[1] + 1:m[2] + 1][::-1].copy()
        done = budget
        assert int(s.stats().moves) == budget and s.stats().converged == 0
        assert (s.tour().astype(np.int64) == prefix).all(), f"after {budget} moves"
    s.run()
    st = s.stats()
    assert (s.tour().astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)
    assert st.converged == 1
    s.close()


def test_mode_r_beyond_one_sm_of_shared_memory_uses_the_per_step_kernel(T, ctx):
    """n = 14 500: the tour records no longer fit one SM's shared memory (k2_two_opt_ref.cu:
    kRefPersistMaxN), so the launch-per-step kernel runs the chain -- same moves, same tour."""
    n = 14500
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    start = O.nn_tour(P, 3)
    want_t, want_st, want_mv = O.two_opt_ref(P, start, log_cap=1 << 16)
    got_t, st, mv = T.Problem.euc2d(ctx, x, y).local_search(T.ALGO_TWO_OPT_REF, start, log_cap=1 << 16)
    assert int(st.launches) > 1000  # one launch per cursor step
    assert [(m[1], m[2]) for m in mv] == [(m[1], m[2]) for m in want_mv]
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)


def test_mode_r_largest_persistent_size(T, ctx):
    """n = 14 000 = kRefPersistMaxN: 224 000 B of records per CTA, a handful of launches."""
    n = 14000
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    start = O.nn_tour(P, 3)
    want_t, want_st, _ = O.two_opt_ref(P, start)
    got_t, st, _ = T.Problem.euc2d(ctx, x, y).local_search(T.ALGO_TWO_OPT_REF, start)
    assert int(st.launches) < 16
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)


# ---- K5 k-NN and the nearest-neighbour constructor ---------------------------------------------

def test_knn_reference_ordered_vector(T, ctx):  # tests/test_kdtree_and_distance_matrix.rs:199-243
    p = T.Problem.euc2d(ctx, [0.0, 1.0, 2.0, 3.0, 10.0, 0.0], [0.0, 0.0, 0.0, 0.0, 0.0, 5.0])
    want = [1, 2, 3, 5, 4]
    for k in range(1, 6):
        assert p.knn(k)[0].tolist() == want[:k]
    assert p.knn(7)[0].tolist() == want + [0xFFFFFFFF] * 2


@pytest.mark.parametrize("k", [1, 3, 5, 8, 12, 32])
def test_knn_matches_oracle(T, ctx, att532, k):
    _, x, y = att532
    got = T.Problem.euc2d(ctx, x, y).knn(k).astype(np.int64)
    assert (got == O.knn(O.Problem(x, y), k)).all()


def test_knn_ties_lower_position_first(T, ctx):
    rng = np.random.default_rng(1)
    x = rng.integers(0, 9, 700).astype(np.float32)
    y = rng.integers(0, 9, 700).astype(np.float32)  # heavy ties and duplicates
    got = T.Problem.euc2d(ctx, x, y).knn(5).astype(np.int64)
    assert (got == O.knn(O.Problem(x, y), 5)).all()


def test_knn_explicit_and_nint(T, ctx, berlin52):
    _, x, y = berlin52
    tri = O.matrix_packed_f32(x, y)
    got = T.Problem.explicit(ctx, tri, 52).knn(5).astype(np.int64)
    assert (got == O.knn(O.Problem(tri=tri, n=52), 5)).all()
    gx, gy = O.gen_grid(800, 800)
    got = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32).knn(5).astype(np.int64)
    assert (got == O.knn(O.Problem(tri=O.matrix_packed_nint(gx, gy), n=800), 5)).all()


def test_nn_tour_goldens_G1_G2(T, ctx, berlin52, att532):
    for (_, x, y), want in ((berlin52, "8980.91797"), (att532, "112099.42188")):
        p = T.Problem.euc2d(ctx, x, y)
        t = p.nn_tour(3)
        assert (t.astype(np.int64) == O.nn_tour(O.Problem(x, y), 3)).all()
        assert f5(p.tour_lengths(t)[0]) == want


@pytest.mark.parametrize("n", [2, 3, 33, 1000, 10000])
def test_nn_tour_matches_oracle(T, ctx, n):
    x, y = O.gen_uniform(n, n)
    t = T.Problem.euc2d(ctx, x, y).nn_tour(3)
    assert (t.astype(np.int64) == O.nn_tour(O.Problem(x, y), 3)).all()


def test_nn_tour_explicit_and_nint(T, ctx, berlin52):
    _, x, y = berlin52
    tri = O.matrix_packed_f32(x, y)
    t = T.Problem.explicit(ctx, tri, 52).nn_tour(3)
    assert (t.astype(np.int64) == O.nn_tour(O.Problem(tri=tri, n=52), 3)).all()
    gx, gy = O.gen_grid(700, 7)
    t = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32).nn_tour(3)
    assert (t.astype(np.int64) == O.nn_tour(O.Problem(tri=O.matrix_packed_nint(gx, gy), n=700), 3)).all()


# ---- K3 Or-opt -------------------------------------------------------------------------------------

def check_or_opt(T, ctx, x, y, start, max_moves=-1):
    P = O.Problem(x, y)
    want_t, want_st, want_mv = O.or_opt(P, start, max_moves=max_moves, log_cap=1 << 16)
    p = T.Problem.euc2d(ctx, x, y)
    got_t, st, mv = p.local_search(T.ALGO_OR_OPT, start, max_moves=max_moves, log_cap=1 << 16)
    assert [m[1:] for m in mv] == [m[1:] for m in want_mv]
    assert [np.float32(m[0]) for m in mv] == [np.float32(m[0]) for m in want_mv]
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)
    return got_t, st


def test_or_opt_golden_G6(T, ctx, berlin52):  # docs/benchmarks.md:48
    _, x, y = berlin52
    P = O.Problem(x, y)
    t, st = check_or_opt(T, ctx, x, y, O.nn_tour(P, 3))
    assert f5(T.Problem.euc2d(ctx, x, y).tour_lengths(t)[0]) == "8097.47607"
    assert (int(st.moves), int(st.passes), int(st.evals)) == (10, 11, 136378)


def test_or_opt_reference_unit_vectors(T, ctx):  # or_opt.rs:246-335
    x, y = [0.0, 1.0, 5.0, 2.0, 3.0], [0.0, 0.0, 5.0, 0.0, 0.0]
    p = T.Problem.euc2d(ctx, x, y)
    s = p.session(T.ALGO_OR_OPT, np.arange(5))
    mv = s.scan()
    assert mv is not None and mv[0] < 0.0 and mv == tuple(
        [np.float32(O.or_opt_find_best(O.Problem(x, y), np.arange(5))[0])] + list(O.or_opt_find_best(O.Problem(x, y), np.arange(5))[1:]))
    s.close()
    psq = T.Problem.euc2d(ctx, [0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0])
    s = psq.session(T.ALGO_OR_OPT, [0, 1, 2, 3])
    assert s.scan() is None  # find_best_move returns None on the optimal square
    s.close()
    t, st, _ = psq.local_search(T.ALGO_OR_OPT, [0, 1, 2, 3])
    assert abs(psq.tour_lengths(t)[0] - 4.0) < 1e-2
    p3 = T.Problem.euc2d(ctx, [0.0, 1.0, 1.0], [0.0, 0.0, 1.0])
    t, st, _ = p3.local_search(T.ALGO_OR_OPT, [2, 0, 1])
    assert t.tolist() == [0, 1, 2]  # n < 4: identity order, seed ignored (or_opt.rs:31-34)


@pytest.mark.parametrize("n", [4, 5, 6, 7, 8, 31, 255, 256, 257, 258, 259, 600])
def test_or_opt_small_and_block_edges(T, ctx, n):
    x, y = O.gen_uniform(n, 300 + n)
    check_or_opt(T, ctx, x, y, O.shuffle_tour(n, n + 2), max_moves=120)


def test_or_opt_1k_nn_start(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    check_or_opt(T, ctx, x, y, O.nn_tour(O.Problem(x, y), 3))


def test_or_opt_ties_on_lattice(T, ctx):
    rng = np.random.default_rng(8)
    x = rng.integers(0, 10, 300).astype(np.float32)
    y = rng.integers(0, 10, 300).astype(np.float32)
    check_or_opt(T, ctx, x, y, O.shuffle_tour(300, 4), max_moves=100)


def test_or_opt_scan_only_10k(T, ctx):
    n = 10000
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    t = O.shuffle_tour(n, 3)
    s = T.Problem.euc2d(ctx, x, y).session(T.ALGO_OR_OPT, t)
    got, want = s.scan(), O.or_opt_find_best(P, t)
    assert got is not None and got[1:] == want[1:] and np.float32(got[0]) == np.float32(want[0])
    s.close()


# ---- matrix-backed path: f32 matrix, TSPLIB nint int32 matrix, EXPLICIT problems -------------------

def read_explicit(path):
    """EXPLICIT TSPLIB -> packed strict lower triangle (tsplib.rs:264-318)."""
    meta, toks, on = {}, [], False
    for line in open(path):
        s = line.strip().upper()
        if s == "EOF":
            break
        if s.endswith("_SECTION"):
            on = s == "EDGE_WEIGHT_SECTION"
            continue
        if on:
            toks += [float(t) for t in s.split()]
        elif ":" in s:
            k, v = s.split(":", 1)
            meta[k.strip()] = v.strip()
    n, fmt = int(meta["DIMENSION"]), meta["EDGE_WEIGHT_FORMAT"]
    if fmt == "FULL_MATRIX":
        m = np.array(toks, dtype=np.float32).reshape(n, n)
    elif fmt == "LOWER_DIAG_ROW":
        m = np.zeros((n, n), dtype=np.float32)
        k = 0
        for i in range(n):
            for j in range(i + 1):
                m[i, j] = toks[k]
                k += 1
    else:
        raise ValueError(fmt)
    return n, np.array([m[i, j] for i in range(1, n) for j in range(i)], dtype=np.float32)


ALGOS = ["best", "cyclic", "ref", "oropt"]


def run_both(T, prob, P, algo, start, path, max_moves=-1):
    code = {"best": T.ALGO_TWO_OPT_BEST, "cyclic": T.ALGO_TWO_OPT_BEST_CYCLIC, "ref": T.ALGO_TWO_OPT_REF,
            "oropt": T.ALGO_OR_OPT}[algo]
    if algo in ("best", "cyclic"):
        want = O.two_opt_best(P, start, cyclic=algo == "cyclic", max_moves=max_moves, nthreads=4, log_cap=1 << 16)
    elif algo == "ref":
        want = O.two_opt_ref(P, start, log_cap=1 << 16)
    else:
        want = O.or_opt(P, start, max_moves=max_moves, log_cap=1 << 16)
    got_t, st, mv = prob.local_search(code, start, path=path, max_moves=-1 if algo == "ref" else max_moves,
                                      log_cap=1 << 16)
    want_t, want_st, want_mv = want
    assert [m[1:] for m in mv] == [m[1:] for m in want_mv], algo
    assert [np.float32(m[0]) for m in mv] == [np.float32(m[0]) for m in want_mv], algo
    assert (got_t.astype(np.int64) == want_t).all(), algo
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals), algo
    return st


@pytest.mark.parametrize("algo", ALGOS)
def test_matrix_path_f32_equals_recompute_and_oracle(T, ctx, berlin52, algo):
    _, x, y = berlin52
    P = O.Problem(x, y)
    st = run_both(T, T.Problem.euc2d(ctx, x, y), P, algo, O.nn_tour(P, 3), T.PATH_MATRIX)
    assert st.path_used == T.PATH_MATRIX


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n", [4, 7, 258, 700])
def test_matrix_path_f32_synthetic(T, ctx, algo, n):
    x, y = O.gen_uniform(n, 400 + n)
    run_both(T, T.Problem.euc2d(ctx, x, y), O.Problem(x, y), algo, O.shuffle_tour(n, n), T.PATH_MATRIX, max_moves=150)


@pytest.mark.parametrize("algo", ALGOS)
def test_nint_i32_matrix_path(T, ctx, algo):
    """TSPLIB nint metric: exact integer deltas, plenty of ties."""
    n = 600
    gx, gy = O.gen_grid(n, 77)
    P = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=n)
    prob = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
    st = run_both(T, prob, P, algo, O.shuffle_tour(n, 5), T.PATH_AUTO, max_moves=200)
    assert st.path_used == T.PATH_MATRIX


def test_mode_r_nint_on_fractional_coordinates(T, ctx):
    """TSPLIB nint of non-integer coordinates (evaluated in double, common.cuh: dist_nint): Mode R's
    persistent kernel recomputes the matrix entries from the coordinates -- same moves as the oracle
    on the packed matrix."""
    n = 400
    x, y = O.gen_uniform(n, 91)
    x, y = (x * np.float32(37.3)).astype(np.float32), (y * np.float32(37.3)).astype(np.float32)
    P = O.Problem(tri=O.matrix_packed_nint(x, y), n=n)
    prob = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
    st = run_both(T, prob, P, "ref", O.shuffle_tour(n, 9), T.PATH_AUTO)
    assert st.path_used == T.PATH_MATRIX and st.converged == 1


def test_mode_r_10k_nint_matches_oracle(T, ctx):
    """BASELINE config 3 (10k cities, int32 TSPLIB-nint matrix), reference-exact first improvement:
    the whole search against the CPU oracle on the packed nint matrix; the Or-opt stage that follows
    still finds the session's Cs records and matrix consistent."""
    n = 10000
    gx, gy = O.gen_grid(n, n)
    P = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=n)
    start = O.nn_tour(P, 3)
    want_t, want_st, want_mv = O.two_opt_ref(P, start, log_cap=1 << 16)
    prob = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
    got_t, st, mv = prob.local_search(T.ALGO_TWO_OPT_REF, start, log_cap=1 << 16)
    assert st.path_used == T.PATH_MATRIX
    assert [m[1:3] for m in mv] == [m[1:3] for m in want_mv]
    assert [np.float32(m[0]) for m in mv] == [np.float32(m[0]) for m in want_mv]
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)
    # a budgeted run stops in the middle of the chain and hands the records back to the matrix kernels
    s = prob.session(T.ALGO_TWO_OPT_REF, start)
    s.run(max_moves=100)
    mid = s.tour()
    s.close()
    want_mid = np.asarray(start, dtype=np.int64).copy()
    for m in want_mv[:100]:
        want_mid[m[1] + 1:m[2] + 1] = want_mid[m[1] + 1:m[2] + 1][::-1].copy()
    assert (mid.astype(np.int64) == want_mid).all()
    want_b = O.two_opt_best(P, mid, max_moves=3, nthreads=8, log_cap=8)
    got_b = prob.local_search(T.ALGO_TWO_OPT_BEST, mid, max_moves=3, log_cap=8)
    assert [m[:3] for m in got_b[2]] == [m[:3] for m in want_b[2]]


def test_nint_berlin52_known_optimum(T, ctx, berlin52, golden_dir):
    ids, x, y = berlin52
    P = O.Problem(tri=O.matrix_packed_nint(x, y), n=52)
    prob = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
    for algo in ALGOS:
        run_both(T, prob, P, algo, O.nn_tour(P, 3), T.PATH_AUTO)


@pytest.mark.parametrize("name", ["gr17.tsp", "ring6_explicit.tsp"])
@pytest.mark.parametrize("algo", ALGOS)
def test_explicit_problems(T, ctx, golden_dir, name, algo):
    n, tri = read_explicit(os.path.join(golden_dir, name))
    P = O.Problem(tri=tri, n=n)
    prob = T.Problem.explicit(ctx, tri, n)
    run_both(T, prob, P, algo, np.arange(n), T.PATH_AUTO)
    run_both(T, prob, P, algo, O.shuffle_tour(n, 3), T.PATH_AUTO)
    with pytest.raises(T.TeelineError):  # no coordinates: the recompute path is illegal
        prob.local_search(T.ALGO_TWO_OPT_BEST, np.arange(n), path=T.PATH_RECOMPUTE)


def test_ring6_explicit_reaches_known_optimum(T, ctx, golden_dir):
    n, tri = read_explicit(os.path.join(golden_dir, "ring6_explicit.tsp"))
    prob = T.Problem.explicit(ctx, tri, n)
    t, _, _ = prob.local_search(T.ALGO_TWO_OPT_BEST_CYCLIC, [0, 2, 4, 1, 3, 5])
    t, _, _ = prob.local_search(T.ALGO_OR_OPT, t)
    assert prob.tour_lengths(t)[0] <= 240.0


def test_matrix_path_repermutation_keeps_results(T, ctx, monkeypatch):
    """Re-laying the matrix in tour order must not change a single move."""
    n = 1500
    x, y = O.gen_uniform(n, 31)
    P = O.Problem(x, y)
    start = O.shuffle_tour(n, 8)
    want_t, want_st, want_mv = O.two_opt_best(P, start, max_moves=90, nthreads=4, log_cap=1 << 12)
    monkeypatch.setenv("TL_REPERMUTE_EVERY", "7")
    prob = T.Problem.euc2d(ctx, x, y)
    got_t, st, mv = prob.local_search(T.ALGO_TWO_OPT_BEST, start, path=T.PATH_MATRIX, max_moves=90, log_cap=1 << 12)
    assert int(st.repermutes) >= 10
    assert [m[1:] for m in mv] == [m[1:] for m in want_mv] and (got_t.astype(np.int64) == want_t).all()


@pytest.mark.parametrize("mode", ["hint", "window"])
@pytest.mark.parametrize("kind", ["f32", "nint"])
def test_matrix_l2_residency_keeps_results(T, ctx, monkeypatch, mode, kind):
    """Keeping part of a larger-than-L2 matrix resident in L2 (eviction hints or an access-policy
    window) is a cache policy only: same moves as the oracle, step after step."""
    n = 9000
    monkeypatch.setenv("TL_MAT_PIN_MB", "48")
    monkeypatch.setenv("TL_MAT_PIN_MODE", mode)
    monkeypatch.setenv("TL_REPERMUTE_EVERY", "5")
    if kind == "f32":
        x, y = O.gen_uniform(n, 77)
        P, dk = O.Problem(x, y), T.DIST_F32_EXACT
    else:
        x, y = O.gen_grid(n, 77)
        P, dk = O.Problem(tri=O.matrix_packed_nint(x, y), n=n), T.DIST_NINT_I32
    start = O.shuffle_tour(n, 5)
    want_t, _, want_mv = O.two_opt_best(P, start, max_moves=12, nthreads=8, log_cap=64)
    prob = T.Problem.euc2d(ctx, x, y, dk)
    got_t, st, mv = prob.local_search(T.ALGO_TWO_OPT_BEST, start, path=T.PATH_MATRIX, max_moves=12, log_cap=64)
    assert int(st.repermutes) >= 2
    assert [m[1:] for m in mv] == [m[1:] for m in want_mv] and (got_t.astype(np.int64) == want_t).all()


def test_matrix_scan_only_10k_f32_and_i32(T, ctx):
    n = 10000
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    t = O.shuffle_tour(n, 2)
    s = T.Problem.euc2d(ctx, x, y).session(T.ALGO_TWO_OPT_BEST, t, T.PATH_MATRIX)
    got, want = s.scan(), O.two_opt_best_scan(P, t, nthreads=8)
    assert got is not None and got[1:3] == want[1:3] and np.float32(got[0]) == np.float32(want[0])
    s.close()
    gx, gy = O.gen_grid(n, n)
    Pi = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=n)
    s = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32).session(T.ALGO_TWO_OPT_BEST, t, T.PATH_AUTO)
    got, want = s.scan(), O.two_opt_best_scan(Pi, t, nthreads=8)
    assert got is not None and got[1:3] == want[1:3] and got[0] == want[0]
    s.close()


# ---- K2 cached Mode B: the same moves as Mode B from cached row minima (k2_two_opt_cached.cu) ------------

def check_cached(T, prob, P, start, path, max_moves=-1):
    """TL_ALGO_TWO_OPT_BEST_CACHED against the oracle's full-scan Mode B: the same move log (positions
    and exact deltas), the same tour, the same move and scan counts; it must have computed fewer pair
    deltas than full scans would."""
    want_t, want_st, want_mv = O.two_opt_best(P, start, max_moves=max_moves, nthreads=4, log_cap=1 << 16)
    got_t, st, mv = prob.local_search(T.ALGO_TWO_OPT_BEST_CACHED, start, path=path, max_moves=max_moves, log_cap=1 << 16)
    assert [(m[1], m[2]) for m in mv] == [(m[1], m[2]) for m in want_mv]
    assert [np.float32(m[0]) for m in mv] == [np.float32(m[0]) for m in want_mv]
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes)) == (want_st.moves, want_st.passes)
    assert bool(st.converged) == (max_moves < 0 or want_st.moves < max_moves)
    assert 0 < int(st.evals) <= want_st.evals * 2 + 8  # row-oriented: never more than two full rescans' worth per scan
    return got_t, st


@pytest.mark.parametrize("n", [4, 5, 6, 7, 13, 33, 257, 600, 1025, 1300])
def test_cached_mode_b_synthetic(T, ctx, n):
    """Shuffled starts: long segments (every row rescanned), short ones, rows whose cached minimum sat on
    a changed column -- on the recompute path and on the f32 matrix."""
    x, y = O.gen_uniform(n, 300 + n)
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    start = O.shuffle_tour(n, n + 1)
    mm = -1 if n <= 600 else 120
    check_cached(T, prob, P, start, T.PATH_RECOMPUTE, max_moves=mm)
    check_cached(T, prob, P, start, T.PATH_MATRIX, max_moves=mm)


def test_cached_mode_b_goldens_ties_and_metrics(T, ctx, berlin52, golden_dir):
    _, x, y = berlin52
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    for start in (O.nn_tour(P, 3), np.arange(52)):
        check_cached(T, prob, P, start, T.PATH_RECOMPUTE)
        check_cached(T, prob, P, start, T.PATH_MATRIX)
    rng = np.random.default_rng(3)
    lx = rng.integers(0, 12, 400).astype(np.float32)
    ly = rng.integers(0, 12, 400).astype(np.float32)  # lattice: many exactly equal deltas
    check_cached(T, T.Problem.euc2d(ctx, lx, ly), O.Problem(lx, ly), O.shuffle_tour(400, 2), T.PATH_RECOMPUTE)
    n = 600  # TSPLIB nint: integer deltas, ties everywhere
    gx, gy = O.gen_grid(n, 77)
    Pn = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=n)
    _, st = check_cached(T, T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32), Pn, O.shuffle_tour(n, 5), T.PATH_AUTO)
    assert st.path_used == T.PATH_MATRIX
    ne, tri = read_explicit(os.path.join(golden_dir, "gr17.tsp"))
    check_cached(T, T.Problem.explicit(ctx, tri, ne), O.Problem(tri=tri, n=ne), np.arange(ne), T.PATH_AUTO)


def test_cached_mode_b_1k_and_budgets(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    nn = O.nn_tour(P, 3)
    t, st = check_cached(T, prob, P, nn, T.PATH_RECOMPUTE)
    assert f5(O.tour_length(P, t)) == "25282.04297" and int(st.moves) == 170
    assert int(st.evals) < 0.2 * 171 * 498501  # the point of the cache
    check_cached(T, prob, P, nn, T.PATH_MATRIX, max_moves=37)
    check_cached(T, prob, P, O.shuffle_tour(1000, 11), T.PATH_RECOMPUTE, max_moves=60)
    for tiny in (2, 3):
        xs, ys = O.gen_uniform(tiny, tiny)
        tt, stt, _ = T.Problem.euc2d(ctx, xs, ys).local_search(T.ALGO_TWO_OPT_BEST_CACHED, np.arange(tiny)[::-1].copy())
        assert tt.tolist() == list(range(tiny))[::-1] and int(stt.moves) == 0 and stt.converged == 1
    s = prob.session(T.ALGO_TWO_OPT_BEST_CACHED, nn)
    with pytest.raises(T.TeelineError):
        s.scan()
    with pytest.raises(T.TeelineError):
        s.set_shard(0, 2)
    s.close()


def test_cached_mode_b_10k_equals_the_full_scan_path(T, ctx):
    """BASELINE config 3 (n = 10 000, nint matrix): the cached search applies the 1508 moves of the
    full-scan search, move for move (the full-scan path itself is checked against the oracle above),
    with a few percent of its pair evaluations."""
    n = 10000
    gx, gy = O.gen_grid(n, n)
    prob = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
    nn = prob.nn_tour(3)
    full_t, full_st, full_mv = prob.local_search(T.ALGO_TWO_OPT_BEST, nn, path=T.PATH_MATRIX, log_cap=1 << 16)
    got_t, st, mv = prob.local_search(T.ALGO_TWO_OPT_BEST_CACHED, nn, path=T.PATH_MATRIX, log_cap=1 << 16)
    assert mv == full_mv and (got_t == full_t).all()
    assert (int(st.moves), int(st.passes)) == (int(full_st.moves), int(full_st.passes)) == (1508, 1509)
    assert int(st.evals) < 0.1 * int(full_st.evals)
    x, y = O.gen_uniform(n, n)
    pf = T.Problem.euc2d(ctx, x, y)
    nnf = pf.nn_tour(3)
    a_t, a_st, a_mv = pf.local_search(T.ALGO_TWO_OPT_BEST, nnf, path=T.PATH_RECOMPUTE, log_cap=1 << 16)
    b_t, b_st, b_mv = pf.local_search(T.ALGO_TWO_OPT_BEST_CACHED, nnf, path=T.PATH_RECOMPUTE, log_cap=1 << 16)
    assert a_mv == b_mv and (a_t == b_t).all() and int(a_st.moves) == int(b_st.moves)


# ---- K2-batch: one CTA per tour, whole search in one launch (multi-start / GA population) ------------

def check_batch(T, ctx, x, y, tours, cyclic=False, max_moves=-1, engines=None):
    """The batched engines -- K2-pop (work items of single scans scheduled over the whole GPU), the
    CTA-per-tour kernel and the cluster-per-tour kernel at every cluster size -- against the oracle's
    complete searches, tour by tour."""
    P = O.Problem(x, y)
    p = T.Problem.euc2d(ctx, x, y)
    algo = T.ALGO_TWO_OPT_BEST_CYCLIC if cyclic else T.ALGO_TWO_OPT_BEST
    want = [O.two_opt_best(P, start, cyclic=cyclic, max_moves=max_moves, nthreads=4) for start in tours]
    moves = sum(w[1].moves for w in want)
    passes = sum(w[1].passes for w in want)
    evals = sum(w[1].evals for w in want)
    saved = {k: os.environ.get(k) for k in ("TL_BATCH_ENGINE", "TL_BATCH_CLUSTER")}
    forced = engines or ([saved["TL_BATCH_ENGINE"]] if saved["TL_BATCH_ENGINE"] else
                         ["pop", "cta", "cluster2", "cluster4", "cluster8", "auto"])
    try:
        for engine in forced:
            os.environ.pop("TL_BATCH_CLUSTER", None)
            os.environ["TL_BATCH_ENGINE"] = "pop" if engine == "pop" else "cta"
            if engine == "cta":
                os.environ["TL_BATCH_CLUSTER"] = "1"
            elif engine.startswith("cluster"):
                os.environ["TL_BATCH_CLUSTER"] = engine[len("cluster"):]
            got, st, lengths = p.two_opt_batch(tours, algo, max_moves=max_moves)
            for b, (want_t, _, _) in enumerate(want):
                assert (got[b].astype(np.int64) == want_t).all(), f"{engine}: tour {b} differs"
                assert bits(lengths[b:b + 1])[0] == bits(np.float32(O.tour_length(P, want_t)))[0], engine
            assert (int(st.moves), int(st.passes), int(st.evals)) == (moves, passes, evals), engine
            # the whole batch is ONE search launch (+ record set-up/extraction for K2-pop) + ONE length launch
            assert int(st.launches) == (4 if engine == "pop" and len(x) >= 4 and max_moves != 0 else 2), engine
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return got, st


@pytest.mark.parametrize("n", [4, 5, 6, 7, 8, 12, 13, 14, 33, 100, 331])
def test_batch_small_n_matches_oracle(T, ctx, n):
    x, y = O.gen_uniform(n, 500 + n)
    tours = np.stack([O.shuffle_tour(n, s) for s in range(1, 20)])
    for cyclic in (False, True):
        check_batch(T, ctx, x, y, tours, cyclic=cyclic)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
def test_batch_every_launch_configuration(T, ctx, monkeypatch, cfg):
    """128 ... 1024 threads per tour (picked by batch size so that a sharded population still fills
    the GPU): the same tours, moves and lengths whatever the CTA size."""
    monkeypatch.setenv("TL_BATCH_CFG", str(cfg))
    n = 700
    x, y = O.gen_uniform(n, 4242)
    tours = np.stack([O.nn_tour(O.Problem(x, y), 3)] + [O.shuffle_tour(n, s) for s in range(1, 6)])
    check_batch(T, ctx, x, y, tours, max_moves=40, engines=["cta"])
    check_batch(T, ctx, x, y, tours[:2], cyclic=True, max_moves=25, engines=["cta"])


@pytest.mark.parametrize("cl", [2, 4, 8])
def test_batch_cluster_per_tour(T, ctx, monkeypatch, cl):
    """A thread-block cluster per tour (DSMEM ticket, candidates exchanged with one cluster barrier per
    step): more tours than resident clusters (every cluster walks several tours), coordinates outside
    the fast-sqrt domain (the IEEE instantiation), a converging run, and a size whose records make
    the kernel fall back to bigger CTAs."""
    engine = [f"cluster{cl}"]
    x, y = O.gen_uniform(300, 77)
    tours = np.stack([O.shuffle_tour(300, s) for s in range(1, 700 // cl)])
    check_batch(T, ctx, x, y, tours, max_moves=6, engines=engine)
    check_batch(T, ctx, x, y, tours[:5], engines=engine)  # to the local optimum
    check_batch(T, ctx, x * np.float32(1e-9), y * np.float32(1e-9), tours[:4], max_moves=30, engines=engine)  # some |c| < 2^-27
    n = 4000  # 65 KB of records: 256 x 4 no longer fits an SM, the plan takes 512 x 2
    x, y = O.gen_uniform(n, 4000)
    tours = np.stack([O.shuffle_tour(n, s) for s in (1, 2, 3)])
    check_batch(T, ctx, x, y, tours, max_moves=8, engines=engine)
    check_batch(T, ctx, x, y, tours[:2], cyclic=True, max_moves=5, engines=engine)


@pytest.mark.parametrize("B", [37, 74, 75, 148, 149, 296, 297, 600])
def test_batch_sizes_around_the_engine_policy(T, ctx, B):
    """Batch sizes on both sides of every threshold of the engine policy (clusters of 8/4/2 1024-thread
    CTAs below S/2 tours, one CTA per tour up to 2 S, clusters of 2 x 256 as the queueing engine above):
    whatever kernel runs, every tour equals the oracle's."""
    n = 200
    x, y = O.gen_uniform(n, 77)
    P, p = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    tours = np.stack([O.shuffle_tour(n, s) for s in range(1, B + 1)])
    got, st, lengths = p.two_opt_batch(tours, max_moves=5)
    assert int(st.moves) == 5 * B and int(st.launches) == 2
    for b in range(B):
        want_t, _, _ = O.two_opt_best(P, tours[b], max_moves=5)
        assert (got[b].astype(np.int64) == want_t).all(), b
    assert (bits(lengths) == bits(p.tour_lengths(got))).all()


def test_batch_berlin52_population(T, ctx, berlin52):
    _, x, y = berlin52
    P = O.Problem(x, y)
    tours = np.stack([O.nn_tour(P, 3)] + [O.shuffle_tour(52, s) for s in range(1, 300)])
    got, st = check_batch(T, ctx, x, y, tours)
    assert st.converged == 1
    # tour 0 is the nn -> 2-opt(best) result of the single-tour path
    t1, _, _ = T.Problem.euc2d(ctx, x, y).local_search(T.ALGO_TWO_OPT_BEST, tours[0])
    assert (got[0] == t1).all()


def test_batch_1k_bounded_and_ties(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    tours = np.stack([O.nn_tour(O.Problem(x, y), 3)] + [O.shuffle_tour(1000, s) for s in range(1, 8)])
    _, st = check_batch(T, ctx, x, y, tours, max_moves=40)
    assert st.converged == 0
    rng = np.random.default_rng(3)
    lx = rng.integers(0, 12, 400).astype(np.float32)
    ly = rng.integers(0, 12, 400).astype(np.float32)  # lattice: many exactly equal deltas
    tours = np.stack([O.shuffle_tour(400, s) for s in range(1, 9)])
    check_batch(T, ctx, lx, ly, tours, max_moves=60)
    check_batch(T, ctx, lx, ly, tours, cyclic=True, max_moves=60)


def test_batch_1k_nn_start_converges_to_survey_probe(T, ctx):
    x, y = O.gen_uniform(1000, 1000)
    P = O.Problem(x, y)
    tours = np.stack([O.nn_tour(P, 3)] * 3)
    got, st = check_batch(T, ctx, x, y, tours)
    assert f5(O.tour_length(P, got[0])) == "25282.04297" and int(st.moves) == 3 * 170


def test_batch_full_size_properties(T, ctx):
    """Config 5 shape (1024 tours x 1000 cities, bounded moves): every result is a permutation,
    no tour got longer, lengths are the exact-order lengths, and a sample matches the oracle."""
    n, B = 1000, 1024
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    p = T.Problem.euc2d(ctx, x, y)
    tours = np.stack([O.shuffle_tour(n, s) for s in range(1, B + 1)])
    before = p.tour_lengths(tours)
    got, st, lengths = p.two_opt_batch(tours, T.ALGO_TWO_OPT_BEST, max_moves=25)
    assert (np.sort(got, axis=1) == np.arange(n, dtype=np.uint32)[None, :]).all()
    assert (lengths < before).all() and int(st.moves) == 25 * B
    assert (bits(lengths) == bits(p.tour_lengths(got))).all()
    for b in (0, 511, 1023):
        want_t, _, _ = O.two_opt_best(P, tours[b], max_moves=25, nthreads=4)
        assert (got[b].astype(np.int64) == want_t).all()


def test_batch_edge_cases(T, ctx):
    x, y = O.gen_uniform(3, 3)
    p = T.Problem.euc2d(ctx, x, y)
    got, st, lengths = p.two_opt_batch(np.array([[2, 0, 1], [0, 1, 2]]), T.ALGO_TWO_OPT_BEST)
    assert got.tolist() == [[2, 0, 1], [0, 1, 2]] and int(st.moves) == 0 and st.converged == 1
    x, y = O.gen_uniform(10, 1)
    p = T.Problem.euc2d(ctx, x, y)
    with pytest.raises(T.TeelineError):
        p.two_opt_batch(np.array([[0, 1, 2, 3, 4, 5, 6, 7, 8, 8]]))
    with pytest.raises(T.TeelineError):
        p.two_opt_batch(np.arange(10)[None, :], T.ALGO_OR_OPT)
    pn = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32)
    with pytest.raises(T.TeelineError):
        pn.two_opt_batch(np.arange(10)[None, :])


@pytest.mark.parametrize("path", ["recompute", "matrix"])
def test_k2_best_ties_when_warps_walk_several_work_items(T, ctx, monkeypatch, path):
    """A tiny grid makes every warp process many work items, out of (i,j) order; ties on a lattice
    must still resolve to the lowest (i,j)."""
    monkeypatch.setenv("TL_MAX_GRID", "2")
    rng = np.random.default_rng(11)
    n = 900
    x = rng.integers(0, 15, n).astype(np.float32)
    y = rng.integers(0, 15, n).astype(np.float32)
    P = O.Problem(x, y)
    start = O.shuffle_tour(n, 6)
    want_t, want_st, want_mv = O.two_opt_best(P, start, max_moves=60, nthreads=4, log_cap=1 << 12)
    p = T.Problem.euc2d(ctx, x, y)
    got_t, st, mv = p.local_search(T.ALGO_TWO_OPT_BEST, start, path=getattr(T, "PATH_" + path.upper()), max_moves=60,
                                   log_cap=1 << 12)
    assert [m[1:3] for m in mv] == [m[1:3] for m in want_mv] and (got_t.astype(np.int64) == want_t).all()


def test_mode_r_10k_full_size_matches_oracle(T, ctx):
    """BASELINE config 3 size, reference-exact mode: 2 636 moves, 7 passes, 349 825 021 evaluations
    (SURVEY.md section 6 probe) -- same tour as the CPU oracle."""
    n = 10000
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    start = O.nn_tour(P, 3)
    want_t, want_st, _ = O.two_opt_ref(P, start)
    got_t, st, _ = T.Problem.euc2d(ctx, x, y).local_search(T.ALGO_TWO_OPT_REF, start)
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (2636, 7, 349825021) == (want_st.moves, want_st.passes, want_st.evals)
    assert f5(O.tour_length(P, got_t)) == "78726.51562"


# ---- K6 3-opt (three_opt.rs) -------------------------------------------------

def check_three_opt(T, ctx, P, prob, start, max_moves=-1, path=None):
    want_t, want_st, want_mv = O.three_opt(P, start, max_moves=max_moves, nthreads=8, log_cap=1 << 12)
    got_t, st, mv = prob.local_search(T.ALGO_THREE_OPT, start, path=T.PATH_AUTO if path is None else path,
                                      max_moves=max_moves, log_cap=1 << 12)
    assert [m[1:] for m in mv] == [m[1:] for m in want_mv]
    assert [np.float32(m[0]) for m in mv] == [np.float32(m[0]) for m in want_mv]
    assert (got_t.astype(np.int64) == want_t).all()
    assert (int(st.moves), int(st.passes), int(st.evals)) == (want_st.moves, want_st.passes, want_st.evals)
    return got_t, st


def test_three_opt_golden_G7(T, ctx, berlin52):  # docs/benchmarks.md:29
    _, x, y = berlin52
    P = O.Problem(x, y)
    prob = T.Problem.euc2d(ctx, x, y)
    t, st = check_three_opt(T, ctx, P, prob, O.nn_tour(P, 3))
    assert f5(prob.tour_lengths(t)[0]) == "7742.64697" and (int(st.moves), int(st.passes)) == (10, 11)
    check_three_opt(T, ctx, P, prob, np.arange(52))
    check_three_opt(T, ctx, P, prob, O.nn_tour(P, 3), path=T.PATH_MATRIX)


def test_three_opt_reference_unit_cases(T, ctx):  # three_opt.rs:263-276 and the n < 4 rule (:25-28)
    p = T.Problem.euc2d(ctx, [0.0, 0.0, 0.0, 1.0, 1.0], [0.0, 0.5, 1.0, 1.0, 0.0])
    t, st, _ = p.local_search(T.ALGO_THREE_OPT, np.arange(5))
    assert t.tolist() == [0, 1, 2, 3, 4] and int(st.moves) == 0 and st.converged == 1
    p3 = T.Problem.euc2d(ctx, [0.0, 1.0, 1.0], [0.0, 0.0, 1.0])
    t, st, _ = p3.local_search(T.ALGO_THREE_OPT, [2, 0, 1])
    assert t.tolist() == [0, 1, 2]
    s = T.Problem.euc2d(ctx, [0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0]).session(T.ALGO_THREE_OPT, [0, 1, 2, 3])
    assert s.scan() is None
    s.close()


@pytest.mark.parametrize("n", [4, 5, 6, 7, 33, 34, 35, 66, 131])
def test_three_opt_small_and_block_edges(T, ctx, n):
    x, y = O.gen_uniform(n, 700 + n)
    check_three_opt(T, ctx, O.Problem(x, y), T.Problem.euc2d(ctx, x, y), O.shuffle_tour(n, n + 3), max_moves=40)


def test_three_opt_ties_nint_and_explicit(T, ctx, golden_dir):
    rng = np.random.default_rng(21)
    x = rng.integers(0, 8, 70).astype(np.float32)
    y = rng.integers(0, 8, 70).astype(np.float32)  # lattice: many equal savings
    check_three_opt(T, ctx, O.Problem(x, y), T.Problem.euc2d(ctx, x, y), O.shuffle_tour(70, 2), max_moves=40)
    gx, gy = O.gen_grid(80, 5)
    Pi = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=80)
    check_three_opt(T, ctx, Pi, T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32), O.shuffle_tour(80, 9), max_moves=30)
    n, tri = read_explicit(os.path.join(golden_dir, "gr17.tsp"))
    check_three_opt(T, ctx, O.Problem(tri=tri, n=n), T.Problem.explicit(ctx, tri, n), np.arange(n))


def test_three_opt_scan_only_600(T, ctx):
    """One full scan of C(600,3) = 35.8 M triples against the threaded oracle."""
    n = 600
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    t = O.shuffle_tour(n, 1)
    s = T.Problem.euc2d(ctx, x, y).session(T.ALGO_THREE_OPT, t)
    got = s.scan()
    want, _ = O.three_opt_find_best(P, t, nthreads=8)
    assert got is not None and got[1:] == want[1:] and np.float32(got[0]) == np.float32(want[0])
    s.close()


# ---- BASELINE-size parity (configs 3, 4, 5 at their full sizes) ------------------------------------------
#
# The oracle cannot run these searches to the end in test time (config 3 Mode B is ~10^11 evaluations),
# so the checks are: single scans compared outright, and for whole searches the logged move sequence is
# replayed on the host and every k-th logged move is RE-DERIVED by an oracle scan of the tour state
# before it; the final tour must be an oracle-verified local optimum.

NCPU = os.cpu_count() or 8


def replay_two_opt(start, moves, upto):
    """Tour state after the first `upto` logged 2-opt moves (swap_2opt, two_opt.rs:69-79)."""
    t = np.asarray(start, dtype=np.int32).copy()
    for (_, i, j, _, _) in moves[:upto]:
        t[i + 1:j + 1] = t[i + 1:j + 1][::-1].copy()
    return t


def test_config3_full_10k_nint_matrix_solve_against_oracle(T, ctx):
    """BASELINE configs[2], the headline workload: n = 10 000 on the 10^6 grid, TSPLIB nint int32 matrix
    in HBM, NN start, Mode B to the local optimum (1508 moves, >= 5 re-lays of the matrix).  Every 50th
    logged move is re-derived by an oracle scan; the final tour is an oracle-verified local optimum;
    the integer tour length drops by exactly the sum of the logged deltas."""
    n = 10000
    gx, gy = O.gen_grid(n, n)
    Pi = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=n)
    start = O.nn_tour(Pi, 3)
    prob = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
    assert (prob.nn_tour(3).astype(np.int64) == start).all()
    got_t, st, mv = prob.local_search(T.ALGO_TWO_OPT_BEST, start, path=T.PATH_MATRIX, log_cap=1 << 12)
    assert st.path_used == T.PATH_MATRIX and st.converged == 1
    assert int(st.moves) == 1508 == len(mv) and int(st.passes) == 1509
    assert int(st.evals) == 1509 * ((n - 3) * (n - 2) // 2)
    assert int(st.repermutes) >= 5
    # the replayed log is the returned tour
    assert (replay_two_opt(start, mv, len(mv)) == got_t.astype(np.int64)).all()
    assert sorted(got_t.tolist()) == list(range(n))
    # sampled re-derivation: state before move k -> the oracle's best move must be move k
    for k in list(range(0, len(mv), 50)) + [len(mv) - 1]:
        want = O.two_opt_best_scan(Pi, replay_two_opt(start, mv, k), nthreads=NCPU)
        assert want is not None and (want[1], want[2]) == (mv[k][1], mv[k][2]) and want[0] == mv[k][0], k
    # local optimum: one more oracle scan finds nothing
    assert O.two_opt_best_scan(Pi, got_t, nthreads=NCPU) is None
    # exact integer bookkeeping
    assert int(O.tour_length(Pi, got_t)) == int(O.tour_length(Pi, start)) + int(sum(m[0] for m in mv))
    assert int(prob.tour_lengths(got_t)[0]) == int(O.tour_length(Pi, got_t))


def test_config3_or_opt_on_the_10k_nint_matrix(T, ctx):
    """Second half of configs[2]: Or-opt on the int32 matrix at n = 10 000 -- scan-level equality from
    two different tours and the first 20 applied moves, each re-derived by the (threaded) oracle scan."""
    n = 10000
    gx, gy = O.gen_grid(n, n)
    Pi = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=n)
    prob = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
    shuffled = O.shuffle_tour(n, 3)
    s = prob.session(T.ALGO_OR_OPT, shuffled, T.PATH_AUTO)
    got, want = s.scan(), O.or_opt_find_best(Pi, shuffled, nthreads=NCPU)
    assert int(s.stats().path_used) == T.PATH_MATRIX
    assert got is not None and got[1:] == want[1:] and got[0] == want[0]
    s.close()
    start = O.nn_tour(Pi, 3)
    got_t, st, mv = prob.local_search(T.ALGO_OR_OPT, start, path=T.PATH_AUTO, max_moves=20, log_cap=64)
    assert int(st.moves) == 20 == len(mv)
    t = start.copy()
    for k, m in enumerate(mv):
        want = O.or_opt_find_best(Pi, t, nthreads=NCPU)
        assert want is not None and want[1:] == m[1:] and want[0] == m[0], k
        t = O.or_opt_apply(t, m[1], m[3], m[2], m[4])
    assert (t == got_t.astype(np.int64)).all()


def test_config3_matrix_tour_equals_recompute_tour_f32(T, ctx):
    """The f32 matrix path and the coordinate-recompute path are bit-identical metrics: the complete
    10k searches must produce the same move log and the same tour (what scripts/converge.py reports)."""
    n = 10000
    x, y = O.gen_uniform(n, n)
    prob = T.Problem.euc2d(ctx, x, y)
    start = prob.nn_tour(3)
    ta, sa, ma = prob.local_search(T.ALGO_TWO_OPT_BEST, start, path=T.PATH_RECOMPUTE, log_cap=1 << 12)
    tb, sb, mb = prob.local_search(T.ALGO_TWO_OPT_BEST, start, path=T.PATH_MATRIX, log_cap=1 << 12)
    assert (ta == tb).all() and ma == mb and int(sa.moves) == int(sb.moves) > 1000
    P = O.Problem(x, y)
    for k in (0, len(ma) // 2, len(ma) - 1):
        want = O.two_opt_best_scan(P, replay_two_opt(start, ma, k), nthreads=NCPU)
        assert (want[1], want[2]) == (ma[k][1], ma[k][2]) and np.float32(want[0]) == np.float32(ma[k][0])
    assert O.two_opt_best_scan(P, ta, nthreads=NCPU) is None


@pytest.mark.parametrize("start_kind", ["nn", "shuffled"])
def test_config4_100k_scan_and_move_log_against_oracle(T, ctx, start_kind):
    """BASELINE configs[3]: n = 100 000, coordinate recompute.  P(100k) = 4 999 750 003 > 2^32, so pair
    counters and work-item arithmetic are exercised beyond 32 bits.  One full scan against the threaded
    oracle scan, then a 20-move search whose every 5th logged move is re-derived by the oracle."""
    n = 100000
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    prob = T.Problem.euc2d(ctx, x, y)
    start = prob.nn_tour(3).astype(np.int32) if start_kind == "nn" else O.shuffle_tour(n, 4)
    s = prob.session(T.ALGO_TWO_OPT_BEST, start, T.PATH_RECOMPUTE)
    got = s.scan()
    want = O.two_opt_best_scan(P, start, nthreads=NCPU)
    assert got is not None and (got[1], got[2]) == (want[1], want[2]) and np.float32(got[0]) == np.float32(want[0])
    s.run(20)
    st, mv, final = s.stats(), s.log(64), s.tour()
    s.close()
    assert int(st.moves) == 20 == len(mv) and int(st.evals) == int(st.passes) * ((n - 3) * (n - 2) // 2)
    assert (mv[0][1], mv[0][2]) == (want[1], want[2])
    for k in (5, 10, 15, 19):
        w = O.two_opt_best_scan(P, replay_two_opt(start, mv, k), nthreads=NCPU)
        assert (w[1], w[2]) == (mv[k][1], mv[k][2]) and np.float32(w[0]) == np.float32(mv[k][0]), k
    assert (replay_two_opt(start, mv, 20) == final.astype(np.int64)).all()


def test_config5_full_population_to_local_optimum(T, ctx):
    """BASELINE configs[4]: 1024 start tours (tour 0 = NN, 1..1023 = splitmix64 shuffles) on the 1k
    instance, every tour to its 2-opt local optimum in one launch.  Sampled tours equal the oracle's
    complete searches; every result is a permutation, a local optimum length-wise (no tour longer than
    its start) and carries its exact-order length."""
    n, B = 1000, 1024
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    p = T.Problem.euc2d(ctx, x, y)
    tours = np.stack([O.nn_tour(P, 3)] + [O.shuffle_tour(n, s) for s in range(1, B)])
    before = p.tour_lengths(tours)
    got, st, lengths = p.two_opt_batch(tours, T.ALGO_TWO_OPT_BEST)
    assert st.converged == 1 and int(st.launches) <= 4
    assert (np.sort(got, axis=1) == np.arange(n, dtype=np.uint32)[None, :]).all()
    assert (lengths <= before).all() and (bits(lengths) == bits(p.tour_lengths(got))).all()
    moves = 0
    for b in (0, 1, 511, 1023):
        want_t, want_st, _ = O.two_opt_best(P, tours[b], nthreads=NCPU)
        assert (got[b].astype(np.int64) == want_t).all(), b
        moves += want_st.moves
    assert f5(lengths[0]) == "25282.04297"  # tour 0 is the survey probe's NN -> Mode B result
    assert int(st.moves) > moves and int(st.passes) == int(st.moves) + B


# ---- round 2 additions: K4 small/large batch kernels, integer nint, argument checks -----------------------

def test_k4_both_kernels_match_oracle(T, ctx):
    """Small batches take the warp-per-tour kernel, large ones the CTA-per-32-tours kernel: both are the
    reference's sequential f32 sum, bit for bit (incl. n not a multiple of 32 and a tour with an
    unknown position)."""
    for n, B in ((1000, 1000), (999, 37), (33, 9), (200, 5000)):
        x, y = O.gen_uniform(n, n + 1)
        P = O.Problem(x, y)
        p = T.Problem.euc2d(ctx, x, y)
        tours = np.stack([O.shuffle_tour(n, s) for s in range(1, min(B, 64) + 1)])
        tours = np.concatenate([tours] * ((B + len(tours) - 1) // len(tours)))[:B].astype(np.uint32)
        tours[B // 2, n // 3] = n + 7  # unknown position -> 0.0 (distance_matrix.rs:221-231)
        want = O.tour_lengths(P, tours[:64]).astype(np.float32)
        got = p.tour_lengths(tours, T.LEN_EXACT)
        assert got[B // 2] == 0.0
        m = np.ones(B, dtype=bool)
        m[B // 2] = False
        ref = np.concatenate([want] * ((B + 63) // 64))[:B]
        assert (bits(got[m]) == bits(ref[m])).all(), (n, B)
        fast = p.tour_lengths(tours, T.LEN_FAST)
        assert np.allclose(fast[m], ref[m], rtol=1e-5)


def test_k1_nint_integer_path_equals_f64_path(T, ctx, monkeypatch):
    """Integer coordinates take the 32-bit integer TSPLIB nint (no FP64 pipe); it must equal both the
    oracle and the library's own f64 path, also at distances 0, 1, r(r+1) boundaries and 1.48e6."""
    rng = np.random.default_rng(11)
    n = 3000
    x = rng.integers(0, 1048577, n).astype(np.float32)
    y = rng.integers(0, 1048577, n).astype(np.float32)
    x[:3], y[:3] = [0, 1048576, 0], [0, 1048576, 0]
    # pairs with d2 = r(r+1) and r(r+1)+1 (the rounding boundary): (0,0)-(r, .) with small offsets
    for k, r in enumerate((1, 2, 3, 7, 20, 99, 1000, 65535)):
        x[10 + 2 * k], y[10 + 2 * k] = 5000.0, 5000.0 + k
        x[11 + 2 * k], y[11 + 2 * k] = 5000.0 + r, 5000.0 + k + (r + 1 if r < 1000 else 1)
    want = O.matrix_packed_nint(x, y)
    got_int = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32).matrix_packed()
    monkeypatch.setenv("TL_NINT_F64", "1")
    got_f64 = T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32).matrix_packed()
    assert (got_int == want).all() and (got_f64 == want).all()


def test_nint_rejects_coordinates_that_would_overflow(T, ctx):
    with pytest.raises(T.TeelineError):
        T.Problem.euc2d(ctx, [0.0, np.inf, 3.0], [0.0, 1.0, 2.0], T.DIST_NINT_I32)
    with pytest.raises(T.TeelineError):
        T.Problem.euc2d(ctx, [0.0, 3.0e7, 3.0], [0.0, 1.0, 2.0], T.DIST_NINT_I32)
    p = T.Problem.euc2d(ctx, [0.0, 1.0e7, 3.0], [0.0, 1.0, 2.0], T.DIST_NINT_I32)  # below 2^24: accepted
    assert p.matrix_packed().tolist() == O.matrix_packed_nint(np.float32([0, 1e7, 3]), np.float32([0, 1, 2])).tolist()


def test_second_large_matrix_session_reuses_the_pool(T, ctx):
    """Two consecutive matrix sessions whose matrices together exceed nothing, but whose first block
    stays cached in the library's pool: the size check must count the cached block as available."""
    n = 6000
    x, y = O.gen_uniform(n, 3)
    p = T.Problem.euc2d(ctx, x, y)
    for _ in range(3):
        s = p.session(T.ALGO_TWO_OPT_BEST, O.shuffle_tour(n, 1), T.PATH_MATRIX)
        assert s.scan() is not None
        s.close()


# ---- K7 Ant System (ant_colony.rs) ------------------------------------------
#
# The reference's RNG is unseeded and it publishes no ACO result (docs/benchmarks.md:52 "to be
# measured"; its own tests only check tour validity, tests/ant_colony_test.rs), so parity here is:
# the CUDA path equals the oracle port -- same Philox stream, same order of f32 additions -- bit for
# bit (best tour and best cost), over many seeds, plus the reference's validity properties.

def check_aco(T, prob, P, seed, init=None, **kw):
    want_t, want_c, _ = O.aco(P, seed, init_tour=init, **kw)
    got_t, got_c, st = prob.aco(seed, init_tour=init, **kw)
    assert sorted(got_t.tolist()) == list(range(P.n))
    assert (got_t.astype(np.int64) == want_t).all(), (seed, kw)
    assert np.float32(got_c) == np.float32(want_c) == np.float32(O.tour_length(P, got_t)), (seed, kw)
    return got_t, got_c, st


def test_aco_berlin52_defaults_equal_the_oracle_port(T, ctx, berlin52):
    _, x, y = berlin52
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    nn = O.nn_tour(P, 3)
    for seed in range(6):
        _, c, st = check_aco(T, prob, P, seed, init=nn)  # AcoOptions::default(): 150 epochs, 25 ants
        assert c <= np.float32(O.tour_length(P, nn)) and int(st.passes) == 150
    for seed in (100, 101):
        check_aco(T, prob, P, seed, init=None, epochs=40)  # no init tour: Philox-shuffled start, flat tau0 = 1


def test_aco_distribution_over_200_seeds(T, ctx, berlin52):
    """200 seeded runs (30 epochs, 10 ants, shuffled warm start as `teeline solve aco` supplies): every
    run equals the oracle port; the spread is sane (all valid, none worse than its start, mean gap to
    the f32 optimum 7544.37 below 25 %)."""
    _, x, y = berlin52
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    costs = []
    for seed in range(200):
        init = O.shuffle_tour(52, 1000 + seed)
        _, c, _ = check_aco(T, prob, P, seed, init=init, epochs=30, num_ants=10)
        assert c <= np.float32(O.tour_length(P, init))
        costs.append(c)
    assert np.mean(costs) < 7544.37 * 1.25 and min(costs) >= 7544.36


@pytest.mark.parametrize("n,epochs,ants", [(3, 5, 4), (4, 5, 3), (257, 6, 7), (300, 10, 25), (1000, 3, 25)])
def test_aco_synthetic_sizes(T, ctx, n, epochs, ants):
    x, y = O.gen_uniform(n, 900 + n)
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    check_aco(T, prob, P, 7, init=O.shuffle_tour(n, 3), epochs=epochs, num_ants=ants)
    check_aco(T, prob, P, 8, init=None, epochs=epochs, num_ants=ants)


def test_aco_exponents_explicit_and_degenerate_weights(T, ctx, golden_dir):
    x, y = O.gen_uniform(120, 5)
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    init = O.shuffle_tour(120, 1)
    for alpha, beta in ((2.0, 3.0), (0.0, 1.0), (3.0, 0.0)):  # exact-product exponents: bit-equal
        check_aco(T, prob, P, 3, init=init, alpha=alpha, beta=beta, epochs=8)
    got_t, got_c, _ = prob.aco(3, init_tour=init, alpha=1.5, beta=2.5, epochs=8)  # powf: validity only
    assert sorted(got_t.tolist()) == list(range(120)) and got_c <= np.float32(O.tour_length(P, init))
    n, tri = read_explicit(os.path.join(golden_dir, "gr17.tsp"))
    check_aco(T, T.Problem.explicit(ctx, tri, n), O.Problem(tri=tri, n=n), 5, init=np.arange(n), epochs=20)
    # distances so large that eta^beta underflows to 0: primary and fallback sums are both 0 and every
    # step takes `fallback.first()`, the first unvisited city (select_next, ant_colony.rs:66-82)
    big = (tri * 0 + 1e25).astype(np.float32)
    pe, Pe = T.Problem.explicit(ctx, big, n), O.Problem(tri=big, n=n)
    t, _, _ = check_aco(T, pe, Pe, 9, init=None, epochs=2, num_ants=3)
    # coincident cities: eta = (1 / MIN_DIST)^2 = 1e12 on their edges
    xd, yd = x.copy(), y.copy()
    xd[:10], yd[:10] = xd[0], yd[0]
    check_aco(T, T.Problem.euc2d(ctx, xd, yd), O.Problem(xd, yd), 4, init=init, epochs=6)


def test_aco_reference_edge_cases_and_option_validation(T, ctx):
    # n <= 2: identity order (ant_colony.rs:107-113)
    p2 = T.Problem.euc2d(ctx, [0.0, 1.0], [0.0, 1.0])
    t, c, _ = p2.aco(1)
    assert t.tolist() == [0, 1] and abs(c - 2 * 2 ** 0.5) < 1e-6
    # epochs = 0: the warm start comes back untouched (test_aco_respects_initial_tour, :358-381)
    p5 = T.Problem.euc2d(ctx, [0.0, 0.0, 0.0, 1.0, 1.0], [0.0, 0.5, 1.0, 1.0, 0.0])
    t, c, _ = p5.aco(1, init_tour=[0, 1, 2, 3, 4], epochs=0)
    assert t.tolist() == [0, 1, 2, 3, 4] and c == 4.0
    for kw, msg in (({"alpha": -1.0}, "alpha must be >= 0"), ({"beta": 7.0}, "beta must be in [0, 6]"),
                    ({"evaporation_rate": 1.0}, "evaporation_rate must be in (0, 1)"),
                    ({"num_ants": 0}, "num_ants must be >= 1"), ({"alpha": float("nan")}, "alpha must be >= 0")):
        with pytest.raises(T.TeelineError) as ei:
            p5.aco(1, **kw)
        assert msg in str(ei.value)
    with pytest.raises(T.TeelineError):
        p5.aco(1, init_tour=[0, 1, 2, 3, 3])


# ---- K8 GA population step (genetic_algorithm.rs) ---------------------------
#
# Same parity notion as K7: the reference's RNG is unseeded, so the CUDA path is compared bit for bit
# (best tour, best length, mutation count) with the oracle port that shares its Philox stream and its
# blocked roulette sums; the crossover operator itself is pinned by the reference's unit vectors
# (tests/test_oracle_goldens.py) and the result spread by docs/benchmarks.md:39,139-141.

def check_ga(T, prob, P, seed, init=None, **kw):
    want_t, want_c, want_st = O.ga(P, seed, init_tour=init, **kw)
    got_t, got_c, st = prob.ga(seed, init_tour=init, **kw)
    assert sorted(got_t.tolist()) == list(range(P.n))
    assert (got_t.astype(np.int64) == want_t).all(), (seed, kw)
    assert np.float32(got_c) == np.float32(want_c) == np.float32(O.tour_length(P, got_t)), (seed, kw)
    assert int(st.moves) == int(want_st.moves) and int(st.evals) == int(want_st.evals)
    assert int(st.passes) == int(want_st.passes)
    return got_t, got_c, st


def test_ga_berlin52_equals_the_oracle_port(T, ctx, berlin52):
    _, x, y = berlin52
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    nn = O.nn_tour(P, 3)
    for seed in range(8):
        check_ga(T, prob, P, seed, init=nn, epochs=400)
        check_ga(T, prob, P, seed, init=None, epochs=200, mutation_probability=0.05)
    _, c, st = check_ga(T, prob, P, 11, init=O.shuffle_tour(52, 12))  # GAOptions::default(): 10 000 epochs
    assert int(st.passes) == 10000 and 7542.0 <= c <= 7542.0 * 1.25


def test_ga_distribution_over_200_seeds(T, ctx, berlin52):
    """200 seeded runs (300 epochs, shuffled warm start as `teeline solve ga` supplies): every run equals
    the oracle port; all valid; the mean improves on the mean start by a wide margin."""
    _, x, y = berlin52
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    costs, starts = [], []
    for seed in range(200):
        init = O.shuffle_tour(52, 2000 + seed)
        _, c, _ = check_ga(T, prob, P, seed, init=init, epochs=300, mutation_probability=0.01)
        costs.append(c)
        starts.append(O.tour_length(P, init))
    assert np.mean(costs) < 0.6 * np.mean(starts) and min(costs) >= 7544.36


@pytest.mark.parametrize("n,epochs", [(2, 5), (3, 5), (4, 20), (5, 20), (7, 30), (8, 30), (255, 12), (257, 12),
                                      (1000, 4)])
def test_ga_synthetic_sizes(T, ctx, n, epochs):
    x, y = O.gen_uniform(n, 700 + n)
    P, prob = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    check_ga(T, prob, P, 7, init=O.shuffle_tour(n, 3), epochs=epochs, mutation_probability=0.2)
    check_ga(T, prob, P, 8, init=None, epochs=epochs, n_elite=0 if n >= 4 else 1)
    check_ga(T, prob, P, 9, init=None, epochs=epochs, n_elite=n, mutation_probability=1.0)  # elites only


def test_ga_explicit_duplicates_and_option_validation(T, ctx, golden_dir):
    n, tri = read_explicit(os.path.join(golden_dir, "gr17.tsp"))
    check_ga(T, T.Problem.explicit(ctx, tri, n), O.Problem(tri=tri, n=n), 5, init=np.arange(n), epochs=200)
    # coincident cities (zero-length edges; equal fitnesses exercise the stable sort's tie order)
    x, y = O.gen_uniform(60, 5)
    x[:20], y[:20] = x[0], y[0]
    check_ga(T, T.Problem.euc2d(ctx, x, y), O.Problem(x, y), 4, init=O.shuffle_tour(60, 1), epochs=150,
             mutation_probability=0.1)
    # all cities coincide: every length is 0 and every fitness is 0 (build_evaluator :117-121)
    z = np.zeros(12, dtype=np.float32)
    check_ga(T, T.Problem.euc2d(ctx, z, z), O.Problem(z, z), 4, init=None, epochs=10)
    p5 = T.Problem.euc2d(ctx, [0.0, 0.0, 0.0, 1.0, 1.0], [0.0, 0.5, 1.0, 1.0, 0.0])
    t, c, _ = p5.ga(1, init_tour=[0, 1, 2, 3, 4], epochs=0)  # test_ga_respects_initial_tour (:353-369)
    assert t.tolist() == [0, 1, 2, 3, 4] and c == 4.0
    for mp in (-0.1, 1.5, float("nan")):
        with pytest.raises(T.TeelineError) as ei:
            p5.ga(1, mutation_probability=mp)
        assert "mutation_probability must be in [0, 1]" in str(ei.value)
    with pytest.raises(T.TeelineError):
        p5.ga(1, init_tour=[0, 1, 2, 3, 3])
    with pytest.raises(T.TeelineError):
        T.Problem.euc2d(ctx, x, y, T.DIST_NINT_I32).ga(1)


def test_batch_beyond_one_cta_takes_the_population_engine(T, ctx):
    """n = 13 000 does not fit one CTA's shared memory: tl_two_opt_batch switches to K2-pop (tour
    records in global memory, work items scheduled over the whole GPU); same moves as the oracle."""
    n = 13000
    x, y = O.gen_uniform(n, n)
    P, p = O.Problem(x, y), T.Problem.euc2d(ctx, x, y)
    tours = np.stack([O.shuffle_tour(n, s) for s in (1, 2, 3)])
    got, st, lengths = p.two_opt_batch(tours, max_moves=4)
    assert int(st.moves) == 12 and int(st.launches) == 4
    for b in range(3):
        want_t, _, _ = O.two_opt_best(P, tours[b], max_moves=4, nthreads=NCPU)
        assert (got[b].astype(np.int64) == want_t).all(), b
    assert (bits(lengths) == bits(p.tour_lengths(got))).all()

"""Pins the CPU oracle against every golden the reference publishes for the hot path
(BASELINE.md "Exact known answers") and against the
reference's inline unit vectors.  CPU only."""
import os

import numpy as np
import pytest

import oracle as O


def f5(v):
    return "%.5f" % v


# ---- G1..G6: published known answers -------------------------------------------------

def test_G1_nn_berlin52(berlin52):  # bench/baseline-solvers.tsv:2-6
    _, x, y = berlin52
    P = O.Problem(x, y)
    assert f5(O.tour_length(P, O.nn_tour(P, 3))) == "8980.91797"


def test_G2_nn_att532(att532):  # bench/baseline-solvers.tsv:12-16 (ATT header => EUC_2D, tsplib.rs:199-202)
    _, x, y = att532
    P = O.Problem(x, y)
    assert f5(O.tour_length(P, O.nn_tour(P, 3))) == "112099.42188"


def test_G3_opt_tour_length(berlin52, golden_dir):  # README.md:365
    ids, x, y = berlin52
    opt = O.read_opt_tour(os.path.join(golden_dir, "berlin52.opt.tour"))
    pos = {int(c): k for k, c in enumerate(ids)}
    P = O.Problem(x, y)
    assert f5(O.tour_length(P, [pos[int(c)] for c in opt])) == "7544.36572"


def test_G4_two_opt_from_identity(berlin52):  # docs/benchmarks.md:28
    _, x, y = berlin52
    P = O.Problem(x, y)
    t, st, _ = O.two_opt_ref(P, np.arange(52))
    assert f5(O.tour_length(P, t)) == "9368.31836"
    assert (st.moves, st.passes, st.evals) == (87, 5, 6125)


def test_G5_nn_then_two_opt(berlin52):  # README.md:385
    _, x, y = berlin52
    P = O.Problem(x, y)
    t, st, _ = O.two_opt_ref(P, O.nn_tour(P, 3))
    assert f5(O.tour_length(P, t)) == "8384.18848"
    assert (st.moves, st.passes, st.evals) == (8, 3, 3675)


def test_G6_nn_then_or_opt(berlin52):  # docs/benchmarks.md:48
    _, x, y = berlin52
    P = O.Problem(x, y)
    t, st, _ = O.or_opt(P, O.nn_tour(P, 3))
    assert f5(O.tour_length(P, t)) == "8097.47607"
    assert (st.moves, st.passes, st.evals) == (10, 11, 136378)


def test_matrix_and_recompute_are_bit_identical(berlin52):
    _, x, y = berlin52
    P = O.Problem(x, y)
    Pm = O.Problem(tri=O.matrix_packed_f32(x, y), n=52)
    nn = O.nn_tour(P, 3)
    assert (O.nn_tour(Pm, 3) == nn).all()
    for fn in (O.two_opt_ref, O.two_opt_best, O.or_opt):
        a, sa, ma = fn(P, nn, log_cap=256)
        b, sb, mb = fn(Pm, nn, log_cap=256)
        assert (a == b).all() and ma == mb and sa.evals == sb.evals


# ---- survey-time synthetic probes -----

def test_synthetic_1k_probe():
    x, y = O.gen_uniform(1000, 1000)
    P = O.Problem(x, y)
    nn = O.nn_tour(P, 3)
    assert f5(O.tour_length(P, nn)) == "29890.67773"
    t, st, _ = O.two_opt_ref(P, nn)
    assert f5(O.tour_length(P, t)) == "25436.69336"
    assert (st.passes, st.moves, st.evals) == (6, 331, 2985018)


@pytest.mark.slow
def test_synthetic_1k_mode_b_probe():
    x, y = O.gen_uniform(1000, 1000)
    P = O.Problem(x, y)
    t, st, _ = O.two_opt_best(P, O.nn_tour(P, 3), nthreads=4)
    assert f5(O.tour_length(P, t)) == "25282.04297"
    assert (st.passes, st.moves, st.evals) == (171, 170, 85073013)


# ---- reference inline unit vectors ---------------------------------------------------

def test_swap_2opt_vectors():  # two_opt.rs:86-98
    assert O.swap_2opt([1, 2, 3, 4], 1, 2).tolist() == [1, 3, 2, 4]
    assert O.swap_2opt([1, 2, 3, 4], 1, 1).tolist() == [1, 2, 3, 4]
    assert O.swap_2opt([1, 2, 3, 4], 2, 1).tolist() == [1, 2, 3, 4]  # from >= to: no-op


TSP5 = ([0.0, 0.0, 0.0, 1.0, 1.0], [0.0, 0.5, 1.0, 1.0, 0.0])


def test_two_opt_tsp5():  # two_opt.rs:100-131
    P = O.Problem(*TSP5)
    t, _, _ = O.two_opt_ref(P, np.arange(5))
    assert t.tolist() == [0, 1, 2, 3, 4]
    assert O.tour_length(P, t) == 4.0
    t, _, _ = O.two_opt_best(P, np.arange(5))
    assert t.tolist() == [0, 1, 2, 3, 4]


def test_two_opt_tiny_n():
    for n in (1, 2, 3):  # n=3: empty loop; n<3 underflows in the reference (documented no-op here)
        P = O.Problem(np.arange(n, dtype=np.float32), np.zeros(n, dtype=np.float32))
        t, st, _ = O.two_opt_ref(P, np.arange(n))
        assert t.tolist() == list(range(n)) and st.moves == 0
        t, st, _ = O.two_opt_best(P, np.arange(n))
        assert t.tolist() == list(range(n)) and st.moves == 0


def test_packed_triangle_vectors():  # distance_matrix.rs:326-349, 371-389
    m = O.matrix_packed_f32([0.0, 0.0, 2.0], [0.0, 1.0, 0.0])
    assert m[0] == 1.0 and m[1] == 2.0 and abs(m[2] - 2.236068) < 1e-6
    m = O.matrix_packed_f32([0.0, 0.0, 2.0, 4.0], [0.0, 1.0, 0.0, 0.0])
    assert len(m) == 6
    P = O.Problem(tri=m, n=4)
    assert O.distance(P, 1, 0) == 1.0 and O.distance(P, 2, 0) == 2.0
    assert O.distance(P, 3, 0) == 4.0 and O.distance(P, 3, 2) == 2.0
    assert abs(O.distance(P, 3, 1) - 4.1231055) < 1e-6
    assert O.distance(P, 1, 3) == O.distance(P, 3, 1) and O.distance(P, 2, 2) == 0.0


def test_tour_length_tsp5_and_short():  # distance_matrix.rs:221-245
    P = O.Problem(*TSP5)
    assert O.tour_length(P, [0, 1, 2, 3, 4]) == 4.0
    assert O.tour_length(P, [3]) == 0.0 and O.tour_length(P, []) == 0.0
    assert O.tour_length(P, [0, 2]) == 2.0  # closing edge + the one window


def test_apply_relocation_vectors():  # or_opt.rs:202-240
    assert O.or_opt_apply([0, 1, 2, 3, 4], 1, 1, 3, False).tolist() == [0, 2, 3, 1, 4]
    assert O.or_opt_apply([0, 1, 2, 3, 4], 3, 1, 0, False).tolist() == [0, 3, 1, 2, 4]
    assert O.or_opt_apply([0, 1, 2, 3, 4], 1, 2, 3, False).tolist() == [0, 3, 1, 2, 4]
    assert O.or_opt_apply([0, 1, 2, 3, 4], 1, 2, 3, True).tolist() == [0, 3, 2, 1, 4]
    assert O.or_opt_apply([0, 1, 2, 3, 4, 5], 1, 3, 4, False).tolist() == [0, 4, 1, 2, 3, 5]


def test_or_opt_find_best_move_vectors():  # or_opt.rs:246-272
    P = O.Problem([0.0, 1.0, 5.0, 2.0, 3.0], [0.0, 0.0, 5.0, 0.0, 0.0])
    mv = O.or_opt_find_best(P, np.arange(5))
    assert mv is not None and mv[0] < 0.0
    Psq = O.Problem([0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0])
    assert O.or_opt_find_best(Psq, [0, 1, 2, 3]) is None


def test_or_opt_solve_vectors():  # or_opt.rs:278-335
    P = O.Problem([0.0, 1.0, 5.0, 2.0, 3.0], [0.0, 0.0, 5.0, 0.0, 0.0])
    t, _, _ = O.or_opt(P, np.arange(5))
    assert O.tour_length(P, t) < O.tour_length(P, np.arange(5))
    Psq = O.Problem([0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0])
    t, _, _ = O.or_opt(Psq, [0, 1, 2, 3])
    assert abs(O.tour_length(Psq, t) - 4.0) < 1e-2
    P3 = O.Problem([0.0, 1.0, 1.0], [0.0, 0.0, 1.0])
    t, st, _ = O.or_opt(P3, [2, 0, 1])
    assert st.moves == 0  # n < 4 guard (the caller returns identity order, or_opt.rs:31-34)
    P6 = O.Problem([0.0, 1.0, 5.0, 2.0, 3.0, 4.0], [0.0, 0.0, 5.0, 0.0, 0.0, 1.0])
    t, _, _ = O.or_opt(P6, np.arange(6))
    assert sorted(t.tolist()) == list(range(6))


def test_knn_ordered_oracle_vector():  # tests/test_kdtree_and_distance_matrix.rs:199-243
    P = O.Problem([0.0, 1.0, 2.0, 3.0, 10.0, 0.0], [0.0, 0.0, 0.0, 0.0, 0.0, 5.0])
    want = [1, 2, 3, 5, 4]
    for k in range(1, 6):
        assert O.knn(P, k)[0].tolist() == want[:k]
    assert O.knn(P, 7)[0].tolist() == want + [-1, -1]  # k > n-1: padded


def test_knn_tie_rule_lower_position_first():  # mod.rs:1839-1858: insert after equal keys, gate d < kth
    P = O.Problem([0.0, 1.0, -1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0, -1.0])
    assert O.knn(P, 2)[0].tolist() == [1, 2]
    assert O.knn(P, 4)[0].tolist() == [1, 2, 3, 4]


CITIES_10 = [(150, 20), (270, 70), (260, 180), (180, 280), (120, 290), (35, 220), (25, 80),
             (80, 25), (155, 155), (90, 140)]  # teeline-web/src/explainers/explainer-cities.ts:15-26


def test_mode_b_cyclic_two_optimal_scenario():  # teeline-web/src/explainers/two-opt.test.ts:76-83
    x = [c[0] for c in CITIES_10]
    y = [c[1] for c in CITIES_10]
    P = O.Problem(x, y)
    assert O.two_opt_best_scan(P, [0, 7, 6, 5, 4, 3, 2, 1, 8, 9], cyclic=True) is None


def test_mode_b_threads_agree():
    x, y = O.gen_uniform(400, 7)
    P = O.Problem(x, y)
    t = O.shuffle_tour(400, 3)
    for cyc in (False, True):
        a = O.two_opt_best_scan(P, t, cyclic=cyc, nthreads=1)
        b = O.two_opt_best_scan(P, t, cyclic=cyc, nthreads=5)
        assert a == b and a is not None


def test_nint_metric_known_answer(berlin52, golden_dir):
    """TSPLIB nint metric: berlin52's optimal tour is 7542 (tests/solvers_integration.rs:8)."""
    ids, x, y = berlin52
    opt = O.read_opt_tour(os.path.join(golden_dir, "berlin52.opt.tour"))
    pos = {int(c): k for k, c in enumerate(ids)}
    Pi = O.Problem(tri=O.matrix_packed_nint(x, y), n=52)
    assert O.tour_length(Pi, [pos[int(c)] for c in opt]) == 7542.0
    assert O.dist_nint(0, 0, 3, 4) == 5 and O.dist_nint(0, 0, 1, 1) == 1 and O.dist_nint(0, 0, 1, 2) == 2


def test_generators_are_deterministic():
    x, y = O.gen_uniform(8, 1000)
    x2, y2 = O.gen_uniform(8, 1000)
    assert (x == x2).all() and (y == y2).all() and x.min() >= 0 and x.max() < 1000
    gx, gy = O.gen_grid(8, 1000)
    assert (gx == np.floor(gx)).all() and gx.max() < 1e6
    t = O.shuffle_tour(100, 5)
    assert sorted(t.tolist()) == list(range(100)) and (t != np.arange(100)).any()


# ---- 3-opt ----------------------------------------------------------

def test_three_opt_golden_G7_berlin52(berlin52):
    """G7: NN -> 3-opt on berlin52 = 7742.65 (docs/benchmarks.md:29); 10 moves, 11 scans."""
    _, x, y = berlin52
    P = O.Problem(x, y)
    t, st, mv = O.three_opt(P, O.nn_tour(P, 3), log_cap=64)
    assert "%.5f" % O.tour_length(P, t) == "7742.64697" and "%.2f" % O.tour_length(P, t) == "7742.65"
    assert (st.moves, st.passes) == (10, 11) and st.evals == 11 * (52 * 51 * 50 // 6 - 50)
    assert sorted(t.tolist()) == list(range(52)) and all(1 <= m[4] <= 7 for m in mv)


def test_three_opt_reference_unit_vectors():
    """apply_3opt vectors (three_opt.rs:284-331) and the seeded-optimal case (:263-276)."""
    base = [0, 1, 2, 3, 4, 5]
    want = {1: [0, 2, 1, 3, 4, 5], 2: [0, 1, 2, 4, 3, 5], 3: [0, 2, 1, 4, 3, 5], 4: [0, 3, 4, 1, 2, 5],
            5: [0, 3, 4, 2, 1, 5], 6: [0, 4, 3, 1, 2, 5], 7: [0, 4, 3, 2, 1, 5]}
    for case, w in want.items():
        assert O.three_opt_apply(base, 0, 2, 4, case).tolist() == w
    x, y = [0.0, 0.0, 0.0, 1.0, 1.0], [0.0, 0.5, 1.0, 1.0, 0.0]
    P = O.Problem(x, y)
    t, st, _ = O.three_opt(P, np.arange(5))
    assert t.tolist() == [0, 1, 2, 3, 4] and st.moves == 0
    assert O.three_opt_find_best(O.Problem([0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0]), np.arange(4))[0] is None


def test_three_opt_threads_agree_and_improves():
    x, y = O.gen_uniform(90, 17)
    P = O.Problem(x, y)
    t = O.shuffle_tour(90, 4)
    a, ev1 = O.three_opt_find_best(P, t, nthreads=1)
    b, ev2 = O.three_opt_find_best(P, t, nthreads=5)
    assert a == b and a is not None and ev1 == ev2 == 90 * 89 * 88 // 6 - 88
    t2 = O.three_opt_apply(t, a[1], a[2], a[3], a[4])
    assert np.float32(O.tour_length(P, t2)) < np.float32(O.tour_length(P, t))


# ---- Ant System port: the pieces the reference pins with unit tests (ant_colony.rs:258-330) ----------

def test_philox_known_answers():
    """Philox4x32-10 against the Random123 known-answer vectors."""
    assert O.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_aco_port_properties(berlin52):
    """Validity and monotonicity the reference's own tests ask for (tests/ant_colony_test.rs:43-95,
    ant_colony.rs:358-381): a permutation, cost == recomputed length, never worse than the warm start,
    epochs = 0 returns the warm start, reproducible for a seed, different across seeds."""
    _, x, y = berlin52
    P = O.Problem(x, y)
    nn = O.nn_tour(P, 3)
    a, ca, _ = O.aco(P, 1, init_tour=nn, epochs=40)
    b, cb, _ = O.aco(P, 1, init_tour=nn, epochs=40)
    c, cc, _ = O.aco(P, 2, init_tour=nn, epochs=40)
    assert sorted(a.tolist()) == list(range(52)) and (a == b).all() and ca == cb
    assert ca <= O.tour_length(P, nn) and abs(O.tour_length(P, a) - ca) < 1e-6
    assert not (a == c).all() or ca != cc
    z, cz, _ = O.aco(P, 1, init_tour=nn, epochs=0)
    assert (z == nn).all() and cz == O.tour_length(P, nn)
    s, cs_, _ = O.aco(P, 3, init_tour=None, epochs=30)  # flat tau0 = 1 path (ant_colony.rs:152-158)
    assert sorted(s.tolist()) == list(range(52)) and cs_ > 0
    sq = O.Problem(np.float32([0, 1]), np.float32([0, 1]))
    t2, c2, _ = O.aco(sq, 1)
    assert t2.tolist() == [0, 1]


# ---- GA population step (genetic_algorithm.rs) ---------------------------------

def test_ordered_crossover_reference_vectors():
    """ordered_crossover_genes pinned by the reference's own unit vectors (genetic_algorithm.rs:458-475)."""
    c1, c2 = O.ox_genes([1, 2, 5, 3, 6, 4], [5, 1, 4, 3, 6, 2], 2, 4)
    assert c1.tolist() == [2, 5, 4, 3, 6, 1] and c2.tolist() == [1, 4, 5, 3, 6, 2]
    c1, c2 = O.ox_genes([9, 8, 4, 5, 6, 7, 1, 3, 2, 0], [8, 7, 1, 2, 3, 0, 9, 5, 4, 6], 3, 5)
    assert c1.tolist() == [5, 6, 7, 2, 3, 0, 1, 9, 8, 4] and c2.tolist() == [2, 3, 0, 5, 6, 7, 9, 4, 8, 1]


def test_ga_port_properties(berlin52):
    """What the reference's tests ask of the GA (genetic_algorithm.rs:353-403): the returned total is the
    real tour length of a permutation; a seeded population starts from the exact warm start (epochs = 0
    returns it when it is the fittest); reproducible for a seed, different across seeds."""
    _, x, y = berlin52
    P = O.Problem(x, y)
    nn = O.nn_tour(P, 3)
    a, ca, st = O.ga(P, 1, init_tour=nn, epochs=300)
    b, cb, _ = O.ga(P, 1, init_tour=nn, epochs=300)
    c, cc, _ = O.ga(P, 2, init_tour=nn, epochs=300)
    assert sorted(a.tolist()) == list(range(52)) and (a == b).all() and ca == cb
    assert abs(O.tour_length(P, a) - ca) < 1e-6 and int(st.passes) == 300 and int(st.evals) == 300 * 2 * 23
    assert not (a == c).all() or ca != cc
    z, cz, _ = O.ga(P, 5, init_tour=nn, epochs=0)  # NN beats every shuffle and mutant: best() is the seed
    assert (z == nn).all() and cz == O.tour_length(P, nn)
    sq = O.Problem(np.float32([0, 0, 1, 1]), np.float32([0, 1, 1, 0]))
    t4, c4, _ = O.ga(sq, 3, epochs=100)  # test_solve_returns_real_tour_length_not_inverted_fitness
    assert sorted(t4.tolist()) == [0, 1, 2, 3] and c4 >= 3.9 and abs(c4 - O.tour_length(sq, t4)) < 0.01


def test_ga_port_distribution_matches_the_published_spread(berlin52):
    """docs/benchmarks.md:39,139-141: berlin52, 10 000 epochs: 8112.46 (+7.5 %) in the table, gap between
    +1 % and +20 % over ten runs.  The seeded port must land in the same band (statistical parity: the
    reference's RNG is unseeded)."""
    _, x, y = berlin52
    P = O.Problem(x, y)
    costs = [O.ga(P, seed, init_tour=O.shuffle_tour(52, seed + 1))[1] for seed in range(4)]
    assert all(7542.0 <= c <= 7542.0 * 1.22 for c in costs), costs

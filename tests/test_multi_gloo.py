"""Multi-process host logic over gloo (world_size 2 and 3, CPU): sharding, gathers and the global
argmin of teeline_b200/multi.py.  The oracle stands in for the CUDA kernels (`solve_fn` is injected),
so these tests cover exactly the code that runs between the kernels on a multi-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle as O
    from teeline_b200 import multi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, B = 60, 11
        x, y = O.gen_uniform(n, 7)
        P = O.Problem(x, y)
        tours = np.stack([O.shuffle_tour(n, s) for s in range(1, B + 1)]).astype(np.uint32)

        def solve_fn(t):  # stand-in for problem.two_opt_batch on this rank's GPU
            outs = [O.two_opt_best(P, row)[0] for row in t]
            return np.stack(outs), np.array([O.tour_length(P, o) for o in outs], dtype=np.float32)

        rng_, out_t, all_len, best, best_tour = multi.sharded_population(tours, solve_fn, dist)
        lens = multi.sharded_tour_lengths(tours, lambda t: O.tour_lengths(P, t).astype(np.float32), dist)

        # triangle sharding: every rank scans a share of the rows, records are merged by (delta, i, j)
        t0 = tours[0].astype(np.int64)
        rows = n - 3
        lo, hi = multi.shard_range(rows, rank, world)
        mine = (np.inf, -1, -1)
        for i in range(lo, hi):
            for j in range(i + 2, n - 1):
                a, b, c, d = t0[i], t0[i + 1], t0[j], t0[j + 1]
                delta = np.float32(np.float32(O.distance(P, a, c)) + np.float32(O.distance(P, b, d))) - \
                    np.float32(np.float32(O.distance(P, a, b)) + np.float32(O.distance(P, c, d)))
                if delta < 0 and (float(delta), i, j) < (float(mine[0]), mine[1] if mine[1] >= 0 else 1 << 30, mine[2]):
                    mine = (float(delta), i, j)
        import torch
        rec = torch.tensor(mine, dtype=torch.float64)
        recs = [torch.empty(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(recs, rec)
        merged = multi.merge_best_records(torch.stack(recs).numpy())
        q.put((rank, rng_, out_t.tolist(), all_len.tolist(), best, best_tour.tolist(), lens.tolist(), merged))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_population_and_triangle_over_gloo(world):
    import torch.multiprocessing as mp

    import oracle as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    n, B = 60, 11
    x, y = O.gen_uniform(n, 7)
    P = O.Problem(x, y)
    tours = np.stack([O.shuffle_tour(n, s) for s in range(1, B + 1)])
    want_t = [O.two_opt_best(P, t)[0] for t in tours]
    want_len = np.array([O.tour_length(P, t) for t in want_t], dtype=np.float32)
    covered = []
    for rank, (lo, hi), out_t, all_len, best, best_tour, lens, merged in results:
        covered += list(range(lo, hi))
        assert [list(map(int, t)) for t in want_t[lo:hi]] == out_t          # each shard solved its own tours
        assert np.array_equal(np.array(all_len, dtype=np.float32), want_len)  # every rank sees every length
        assert best == int(np.argmin(want_len)) and best_tour == list(map(int, want_t[best]))
        assert np.array_equal(np.array(lens, dtype=np.float32), O.tour_lengths(P, tours).astype(np.float32))
        mv = O.two_opt_best_scan(P, tours[0])
        assert merged is not None and (merged[1], merged[2]) == (mv[1], mv[2]) and np.float32(merged[0]) == np.float32(mv[0])
    assert covered == list(range(B))  # shards tile the population exactly once


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from teeline_b200 import multi
    for total in (0, 1, 7, 1024, 1000):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multi.shard_range(10, 2, 2)

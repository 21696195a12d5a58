"""Worker of tests/test_gpu_multi.py: run under torchrun, one process per GPU.

Every assertion is against the CPU ORACLE (rank 0 computes it, the expectation is broadcast), never
against an unsharded GPU run: the sharded triangle (config 4) through both transports -- the
in-kernel peer-mailbox exchange and the ncclAllGather + apply-kernel path -- on the recompute and
the int32 matrix paths, and the sharded population (config 5)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle as O  # noqa: E402
import teeline_b200 as T  # noqa: E402
from teeline_b200 import multi  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NCPU = max(1, (os.cpu_count() or 8))


def bcast(obj):
    box = [obj]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def replay(start, moves, upto):
    t = np.asarray(start, dtype=np.int32).copy()
    for (_, i, j, _, _) in moves[:upto]:
        t[i + 1:j + 1] = t[i + 1:j + 1][::-1].copy()
    return t


def all_equal(obj):
    """Every rank holds the same python object."""
    objs = [None] * world
    dist.all_gather_object(objs, obj)
    return all(o == objs[0] for o in objs)


def check_sharded(ctx, prob, P, start, moves, path, label, sample_every, exact_int=False):
    s = prob.session(T.ALGO_TWO_OPT_BEST, start, path)
    s.set_shard(rank, world)
    first = s.scan()  # all-gather + host reduction: the global best from the start tour
    s.run(moves)
    st, log, tour = s.stats(), s.log(moves + 8), s.tour()
    s.close()
    assert int(st.moves) == moves == len(log), (label, int(st.moves), len(log))
    assert all_equal((log, tour.tolist())), f"{label}: ranks disagree"
    ks = sorted(set(list(range(0, moves, sample_every)) + [moves - 1]))
    want = bcast([O.two_opt_best_scan(P, replay(start, log, k), nthreads=NCPU) for k in ks] if rank == 0 else None)
    for k, w in zip(ks, want):
        assert w is not None and (w[1], w[2]) == (log[k][1], log[k][2]), (label, k, w, log[k])
        assert (w[0] == log[k][0]) if exact_int else (np.float32(w[0]) == np.float32(log[k][0])), (label, k)
    assert first is not None and first[1:3] == log[0][1:3], label
    assert (replay(start, log, moves) == tour.astype(np.int64)).all(), label
    return log


def main():
    ctx = T.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    multi.attach_nccl(ctx, dist)
    report = {}
    for transport in ("mailbox", "nccl"):
        if transport == "nccl":
            os.environ["TL_SHARD_TRANSPORT"] = "nccl"
        else:
            os.environ.pop("TL_SHARD_TRANSPORT", None)
        # recompute, 20k shuffled tour: every 6th of 40 moves re-derived by the oracle
        n = 20000
        x, y = O.gen_uniform(n, n)
        prob, P = T.Problem.euc2d(ctx, x, y), O.Problem(x, y)
        a = check_sharded(ctx, prob, P, O.shuffle_tour(n, 3), 40, T.PATH_RECOMPUTE, f"{transport}/recompute20k", 6)
        report[f"{transport}_20k"] = a[:3]
        prob.close()
        # int32 matrix path, 4000 cities on the grid, ties possible: every move of 30 checked
        n = 4000
        gx, gy = O.gen_grid(n, 77)
        prob = T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32)
        Pi = O.Problem(tri=O.matrix_packed_nint(gx, gy), n=n)
        check_sharded(ctx, prob, Pi, O.shuffle_tour(n, 5), 30, T.PATH_MATRIX, f"{transport}/nint4k", 1, exact_int=True)
        prob.close()
    os.environ.pop("TL_SHARD_TRANSPORT", None)
    assert report["mailbox_20k"] == report["nccl_20k"]

    # config 4 at full size: n = 100 000 (P > 2^32), NN start, 12 moves, every 4th re-derived
    n = 100000
    x, y = O.gen_uniform(n, n)
    prob, P = T.Problem.euc2d(ctx, x, y), O.Problem(x, y)
    start = prob.nn_tour(3).astype(np.int32)
    assert all_equal(start.tolist())
    check_sharded(ctx, prob, P, start, 12, T.PATH_RECOMPUTE, "mailbox/recompute100k", 4)
    prob.close()

    # sharded search to convergence (small instance): same final tour as the oracle's complete search
    n = 1500
    x, y = O.gen_uniform(n, 31)
    prob, P = T.Problem.euc2d(ctx, x, y), O.Problem(x, y)
    start = O.nn_tour(P, 3)
    s = prob.session(T.ALGO_TWO_OPT_BEST, start, T.PATH_RECOMPUTE)
    s.set_shard(rank, world)
    s.run(-1)
    st, tour = s.stats(), s.tour()
    s.close()
    want = bcast(O.two_opt_best(P, start, nthreads=NCPU)[0].tolist() if rank == 0 else None)
    assert st.converged == 1 and tour.tolist() == want

    # config 5: population sharded by tour index, results gathered; sampled tours against the oracle
    n, B = 1000, 64
    x, y = O.gen_uniform(n, n)
    prob, P = T.Problem.euc2d(ctx, x, y), O.Problem(x, y)
    tours = np.stack([O.nn_tour(P, 3)] + [O.shuffle_tour(n, s) for s in range(1, B)]).astype(np.uint32)
    (lo, hi), mine, all_len, best, best_tour = multi.sharded_population(
        tours, lambda t: prob.two_opt_batch(t)[::2], dist)
    picks = [0, B // 2 - 1, B // 2, B - 1]
    want = bcast({b: O.two_opt_best(P, tours[b], nthreads=NCPU)[0].tolist() for b in picks} if rank == 0 else None)
    for b in picks:
        if lo <= b < hi:
            assert mine[b - lo].tolist() == want[b], b
        assert np.float32(all_len[b]) == np.float32(O.tour_length(P, np.array(want[b])))
    assert best == int(np.argmin(all_len)) and all_equal((best, best_tour.tolist(), all_len.tolist()))
    prob.close()

    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    ctx.close()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world}", flush=True)


if __name__ == "__main__":
    main()

"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

Two patterns, both from SURVEY.md section 8(e):

* population work (BASELINE config 5: B independent start tours / GA population): tours are
  sharded by index, every rank runs its shard with NO data-path collective, and only the result
  summary crosses ranks -- one all-gather of (length) per tour and one broadcast of the best tour;
* one huge instance (config 4): every rank holds a replica of the tour, `Session.set_shard`
  gives it an equal share of the (i,j) triangle and the library exchanges the per-rank best
  records with one ncclAllGather per scan (see csrc/session.cu).

The compute itself is injected (`solve_fn`), so the same code runs over NCCL on GPUs and over
gloo in the CPU tests (tests/test_multi_gloo.py), where the oracle stands in for the kernels.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total` units: the first `total % world` ranks get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist_device(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def attach_nccl(ctx, dist) -> None:
    """Give the library its own NCCL communicator: rank 0 creates the unique id, torch.distributed
    broadcasts the 128 bytes, every rank attaches (include/teeline_cuda.h: tl_ctx_attach_nccl)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _dist_device(dist)
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = torch.frombuffer(bytearray(type(ctx).nccl_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(buf, src=0)
    ctx.attach_nccl(bytes(buf.cpu().numpy().tobytes()), rank, world)


def gather_lengths(my_lengths: np.ndarray, total: int, dist) -> np.ndarray:
    """All-gather the per-tour results of every rank's shard into one array of `total` entries."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _dist_device(dist)
    width = (total + world - 1) // world  # equal-size slots so the all-gather is symmetric
    slot = torch.full((width,), float("inf"), dtype=torch.float32, device=dev)
    slot[: len(my_lengths)] = torch.from_numpy(np.ascontiguousarray(my_lengths, dtype=np.float32)).to(dev)
    slots = [torch.empty(width, dtype=torch.float32, device=dev) for _ in range(world)]
    dist.all_gather(slots, slot)
    parts = torch.stack(slots).cpu().numpy()
    full = np.empty(total, dtype=np.float32)
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        full[lo:hi] = parts[r, : hi - lo]
    return full


def sharded_population(tours: np.ndarray, solve_fn: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]], dist):
    """Run `solve_fn` (e.g. `lambda t: problem.two_opt_batch(t)[::2]`) on this rank's shard of
    `tours` (batch x n) and agree on the global best.

    Returns (my_range, my_tours_out, all_lengths, best_index, best_tour).  Ties between equal
    lengths resolve to the lowest tour index, so every rank picks the same winner."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    total, n = tours.shape
    lo, hi = shard_range(total, rank, world)
    mine = np.ascontiguousarray(tours[lo:hi])
    if hi > lo:
        out_tours, lengths = solve_fn(mine)
    else:
        out_tours, lengths = mine, np.empty(0, dtype=np.float32)
    all_lengths = gather_lengths(np.asarray(lengths, dtype=np.float32), total, dist)
    best = int(np.argmin(all_lengths))  # numpy argmin returns the first minimum: lowest index wins ties
    owner = next(r for r in range(world) if shard_range(total, r, world)[0] <= best < shard_range(total, r, world)[1])
    dev = _dist_device(dist)
    buf = torch.zeros(n, dtype=torch.int64, device=dev)
    if rank == owner:
        buf = torch.from_numpy(np.asarray(out_tours[best - lo], dtype=np.int64)).to(dev)
    dist.broadcast(buf, src=owner)
    return (lo, hi), out_tours, all_lengths, best, buf.cpu().numpy().astype(np.uint32)


def sharded_tour_lengths(tours: np.ndarray, length_fn: Callable[[np.ndarray], np.ndarray], dist) -> np.ndarray:
    """Population fitness (K4) sharded by tour index; every rank gets all lengths."""
    total = tours.shape[0]
    lo, hi = shard_range(total, dist.get_rank(), dist.get_world_size())
    mine = length_fn(np.ascontiguousarray(tours[lo:hi])) if hi > lo else np.empty(0, dtype=np.float32)
    return gather_lengths(mine, total, dist)


def merge_best_records(records: np.ndarray):
    """Reference implementation of the cross-rank argmin the library performs on the device after
    its all-gather: `records` is (k, 3) rows of (delta, i, j) with i < 0 meaning "no candidate";
    returns the lexicographic minimum (delta, i, j) or None.  Used by the tests to check that
    sharding the triangle cannot change the selected move."""
    best = None
    for d, i, j in records:
        if i < 0:
            continue
        key = (float(d), int(i), int(j))
        if best is None or key < best:
            best = key
    return best

"""Multi-GPU parity + scaling check (run with torchrun on N GPUs of one box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      scripts/multi_gpu_check.py [out.json]
Config 4: one instance, the (i,j) triangle sharded over ranks, one ncclAllGather of best records per
scan; every rank must end with the same tour as an unsharded run.  Config 5: 1024 start tours sharded
by index, no data-path collective, best tour agreed at the end."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import teeline_b200 as T  # noqa: E402
from teeline_b200 import multi  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = T.Context(local, stream=torch.cuda.current_stream().cuda_stream)
multi.attach_nccl(ctx, dist)
out = {"world": world}
# NCCL connects lazily on the first collective: do that outside every timed region
_w = torch.zeros(1, device="cuda")
dist.all_reduce(_w)


def tmax(v):
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- config 4: sharded triangle ---------------------------------------------------------------------
for n, moves in ((20000, 60), (100000, 20)):
    x, y = bench.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    start = bench.shuffle_tour(n, 3) if n <= 20000 else p.nn_tour(3)
    ref = p.session(T.ALGO_TWO_OPT_BEST, start, T.PATH_RECOMPUTE)          # unsharded replica
    ref.run(moves)
    want, want_log = ref.tour(), ref.log(moves)
    ref_scan_ms = ref.time_scans(5)
    ref.close()
    w = p.session(T.ALGO_TWO_OPT_BEST, start, T.PATH_RECOMPUTE)  # warms the library's own communicator
    w.set_shard(rank, world)
    w.run(2)
    w.close()
    s = p.session(T.ALGO_TWO_OPT_BEST, start, T.PATH_RECOMPUTE)
    s.set_shard(rank, world)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run(moves)
    e1.record()
    torch.cuda.synchronize()
    got, got_log = s.tour(), s.log(moves)
    same = bool((got == want).all()) and [m[1:3] for m in got_log] == [m[1:3] for m in want_log] and \
        [np.float32(m[0]) for m in got_log] == [np.float32(m[0]) for m in want_log]
    flags = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    shard_scan_ms = tmax(s.time_scans(5))  # scan of this rank's shard + the all-gather, max over ranks
    step_ms = tmax(e0.elapsed_time(e1)) / moves
    pairs = (n - 3) * (n - 2) // 2
    out[f"config4_n{n}"] = {"moves": moves, "identical_to_unsharded_on_every_rank": bool(flags.item()),
                            "unsharded_scan_ms": ref_scan_ms, "sharded_scan_plus_allgather_ms": shard_scan_ms,
                            "sharded_step_ms": step_ms, "scan_speedup": ref_scan_ms / shard_scan_ms,
                            "moves_per_s_sharded_scan": pairs / (shard_scan_ms * 1e-3)}
    s.close()
    p.close()

# ---- config 5: sharded population ----------------------------------------------------------------------
n, B = 1000, 1024
x, y = bench.gen_uniform(n, n)
p = T.Problem.euc2d(ctx, x, y)
tours = np.stack([p.nn_tour(3)] + [bench.shuffle_tour(n, s) for s in range(1, B)])
p.two_opt_batch(tours[:4], max_moves=2)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
(lo, hi), mine, all_len, best, best_tour = multi.sharded_population(tours, lambda t: p.two_opt_batch(t)[::2], dist)
torch.cuda.synchronize()
dt = tmax(time.perf_counter() - t0)
# weak scaling of the same case: every rank solves its OWN 1024-tour population (different shuffles),
# no data-path collective; the time is the slowest rank's wall clock
wtours = np.stack([bench.shuffle_tour(n, 100000 * (rank + 1) + s) for s in range(B)])
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
_, wst, _ = p.two_opt_batch(wtours)
torch.cuda.synchronize()
wdt = tmax(time.perf_counter() - t0)
wev = torch.tensor([float(wst.evals)], device="cuda", dtype=torch.float64)
dist.all_reduce(wev)
out["config5_weak_1024_per_gpu"] = {"tours_total": B * world, "slowest_rank_wall_s": wdt, "evals_total": float(wev.item()),
                                    "evals_per_s_aggregate": float(wev.item()) / wdt}
single = None
if rank == 0:
    t0 = time.perf_counter()
    got, st, lengths = p.two_opt_batch(tours)
    single = time.perf_counter() - t0
    ok = bool((np.float32(all_len) == lengths).all()) and best == int(np.argmin(lengths)) and bool((best_tour == got[best]).all())
    out["config5_1024x1000"] = {"sharded_wall_s": dt, "single_gpu_wall_s": single, "speedup": single / dt,
                                "identical_lengths_and_best_tour": ok, "best_length": float(lengths.min()),
                                "evals": int(st.evals), "evals_per_s_sharded": int(st.evals) / dt}
    print(json.dumps(out), flush=True)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", f"multi_gpu_{world}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
dist.barrier()
dist.destroy_process_group()

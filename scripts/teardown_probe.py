"""Teardown-order probe for a context with NCCL + peer mailboxes attached (run under torchrun):
   teardown_probe.py <variant>   variant: close_first | close_after | no_close"""
import faulthandler
import os
import sys

faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import teeline_b200 as T  # noqa: E402
from teeline_b200 import multi  # noqa: E402

variant = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = T.Context(local, stream=torch.cuda.current_stream().cuda_stream)
multi.attach_nccl(ctx, dist)
x, y = bench.gen_uniform(5000, 5)
p = T.Problem.euc2d(ctx, x, y)
s = p.session(T.ALGO_TWO_OPT_BEST, bench.shuffle_tour(5000, 2), T.PATH_RECOMPUTE)
s.set_shard(rank, world)
s.run(10)
print(f"[{variant}] rank {rank} moves {int(s.stats().moves)}", flush=True)
s.close()
p.close()
dist.barrier()
if variant == "close_first":
    ctx.close()
    print(f"[{variant}] rank {rank} ctx closed", flush=True)
    dist.destroy_process_group()
elif variant == "close_after":
    dist.destroy_process_group()
    ctx.close()
    print(f"[{variant}] rank {rank} ctx closed", flush=True)
else:
    dist.destroy_process_group()
print(f"[{variant}] rank {rank} done", flush=True)

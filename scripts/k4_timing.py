"""Warm K4 timing (GPU box): TL_K4_TIMING=1 python scripts/k4_timing.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, teeline_b200 as T
ctx = T.Context(0)
rng = np.random.default_rng(1)
for n, B in ((1000, 1000), (1000, 1024), (1000, 4096), (1000, 65536), (5000, 8192), (20000, 2048)):
    x, y = bench.gen_uniform(n, n)
    p = T.Problem.euc2d(ctx, x, y)
    base = np.stack([rng.permutation(n) for _ in range(64)]).astype(np.uint32)
    tours = np.tile(base, ((B + 63) // 64, 1))[:B]
    p.tour_lengths(tours, T.LEN_EXACT)
    p.tour_lengths(tours, T.LEN_FAST)

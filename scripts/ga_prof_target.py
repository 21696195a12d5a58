"""Profiling target for K8 (run under ncu): `ga_prof_target.py <n> <epochs>`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
n, epochs = int(sys.argv[1]), int(sys.argv[2])
ctx = T.Context(0)
x, y = bench.gen_uniform(n, n)
p = T.Problem.euc2d(ctx, x, y)
t, c, st = p.ga(1, init_tour=bench.shuffle_tour(n, 5), epochs=epochs)
print({"best": c, "device_ms": st.device_ms, "launches": int(st.launches)})

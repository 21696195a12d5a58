#!/usr/bin/env python
"""bench.py -- headline benchmark: 2-opt moves evaluated per second at n = 10 000.

Contract (see the task prompt, section 4): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[2]): 10 000 uniform-random EUC_2D cities
(splitmix64, seed 10000, [0,1000)^2 on a 2^-24 grid), nearest-neighbour start tour,
best-improvement 2-opt ("Mode B").  A STEP is one full scan of the
P(n) = (n-3)(n-2)/2 = 49 975 003 candidate moves plus the application of the best one.
`value` = moves evaluated per second with everything resident in HBM.
`e2e`   = the same metric through the C ABI with HOST buffers: each e2e step is one
          tl_problem_create_euc2d + tl_local_search(max_moves = E) + tour read-back.
With N > 1 every rank runs an independent replica from its own start tour
(multi-start; no data-path collective) -> weak scaling.

`--impl reference` times the CPU oracle port of the same step (the reference itself is
Rust and cannot be built in this image) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "2-opt moves evaluated/sec at n=10k (best-improvement scan+apply)"
UNIT = "moves/s"

WORKLOADS = {
    # name: (n, seed)
    "n10k": (10_000, 10_000),
    "n1k": (1_000, 1_000),
    "n100k": (100_000, 100_000),
}


# ---- synthetic instance (same stream as oracle/teeline_oracle.c: tlo_gen_uniform) ----------------

def splitmix64_stream(seed: int, count: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def gen_uniform(n: int, seed: int):
    u = (splitmix64_stream(seed, 2 * n) >> np.uint64(40)).astype(np.float32)
    scale = np.float32(1000.0) / np.float32(16777216.0)
    xy = (u * scale).astype(np.float32)
    return np.ascontiguousarray(xy[0::2]), np.ascontiguousarray(xy[1::2])


def gen_grid(n: int, seed: int):
    """Integer grid floor(u * 10^6) from the same stream (NINT_I32 workloads; oracle: tlo_gen_grid)."""
    u = (splitmix64_stream(seed, 2 * n) >> np.uint64(40)).astype(np.float64)
    g = np.floor(u / 16777216.0 * 1.0e6).astype(np.float32)
    return np.ascontiguousarray(g[0::2]), np.ascontiguousarray(g[1::2])


def shuffle_tour(n: int, seed: int) -> np.ndarray:
    """Fisher-Yates driven by splitmix64(seed) (oracle: tlo_shuffle_tour)."""
    r = splitmix64_stream(seed, n - 1)
    t = np.arange(n, dtype=np.uint32)
    k = 0
    for i in range(n - 1, 0, -1):
        j = int(r[k] % np.uint64(i + 1))
        k += 1
        t[i], t[j] = t[j], t[i]
    return t


def pairs_per_scan(n: int) -> int:
    return (n - 3) * (n - 2) // 2


# ---- clocks sampling ------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline (the oracle; never on the product path) ----------------------------------------------

def cpu_scan_rate(n: int, seed: int, threads: int, budget_s: float):
    """Times full Mode B scans of the oracle on the host cores; returns (moves/s, scans, seconds)."""
    import oracle as O
    x, y = O.gen_uniform(n, seed)
    P = O.Problem(x, y)
    tour = O.nn_tour(P, 3)
    O.two_opt_best_scan(P, tour, nthreads=threads)  # warm
    t0 = time.perf_counter()
    scans = 0
    while True:
        O.two_opt_best_scan(P, tour, nthreads=threads)
        scans += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return scans * pairs_per_scan(n) / dt, scans, dt


def run_reference(args, rank: int, world: int):
    """--impl reference: the CPU port of the same step (scan + apply), all host threads."""
    if rank != 0:
        return
    import oracle as O
    n, seed = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    x, y = O.gen_uniform(n, seed)
    P = O.Problem(x, y)
    tour = O.nn_tour(P, 3).copy()
    budget = 150.0
    done, t_all = 0, 0.0
    t_start = time.perf_counter()
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        mv = O.two_opt_best_scan(P, tour, nthreads=threads)
        if mv is not None:
            tour = O.swap_2opt(tour, mv[1] + 1, mv[2])
        dt = time.perf_counter() - t0
        if k >= args.warmup:
            done += 1
            t_all += dt
        if time.perf_counter() - t_start > budget and done >= 1:
            break
    value = done * pairs_per_scan(n) / t_all
    sample = (f"{done} of {args.steps} requested full scan+apply steps of the n={n} workload "
              f"(time-boxed to {budget:.0f} s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: n={n} uniform EUC_2D seed {seed}, NN start, Mode B scan+apply",
                   "note": "CPU oracle port of the reference semantics (Rust reference not buildable here: no cargo)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- product arm -------------------------------------------------------------------------------------

def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import teeline_b200 as T

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n, seed = WORKLOADS[args.workload]
    P = pairs_per_scan(n)
    x, y = gen_uniform(n, seed)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = T.Context(local_rank, stream=stream)
    prob = T.Problem.euc2d(ctx, x, y)
    # start tour: rank 0 = nearest neighbour (the reference's `nn,2opt` pipeline); other ranks
    # = independent multi-start tours
    start = prob.nn_tour(3) if rank == 0 else shuffle_tour(n, rank)
    path = {"recompute": T.PATH_RECOMPUTE, "matrix": T.PATH_MATRIX, "auto": T.PATH_AUTO}[args.path]
    sess = prob.session(T.ALGO_TWO_OPT_BEST, start, path)

    # --- device-resident timing: W warm-up steps, then exactly K steps
    sess.enqueue(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    sess.enqueue(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    st = sess.stats()
    real_steps = int(st.moves)
    if real_steps < args.warmup + args.steps:
        raise SystemExit(f"bench invalid: the tour converged after {real_steps} moves, fewer than "
                         f"warmup+steps={args.warmup + args.steps}; lower --steps")
    if dist is not None:
        tms = torch.tensor([ms], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * args.steps * P / (ms * 1e-3)

    # --- dominant kernel: average launch duration of the scan kernel (CUDA events, same stream)
    scan_ms = sess.time_scans(max(10, min(args.steps, 200)))
    stats_path = int(st.path_used)

    # --- the matrix-backed paths on the same workload (BASELINE configs[2]: "int32 matrix-backed"):
    #     f32 matrix of the same instance, and the TSPLIB nint int32 matrix of the integer-grid instance
    other_paths = {}
    if rank == 0 and args.workload != "n100k":
        for label, kind in (("matrix_f32", T.DIST_F32_EXACT), ("matrix_nint_i32", T.DIST_NINT_I32)):
            if stats_path == T.PATH_MATRIX and kind == T.DIST_F32_EXACT:
                continue
            gx, gy = (x, y) if kind == T.DIST_F32_EXACT else gen_grid(n, seed)
            p2 = T.Problem.euc2d(ctx, gx, gy, kind)
            s2 = p2.session(T.ALGO_TWO_OPT_BEST, p2.nn_tour(3), T.PATH_MATRIX)
            s2.enqueue(args.warmup)
            torch.cuda.synchronize()
            k = min(args.steps, 100)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            s2.enqueue(k)
            a1.record()
            torch.cuda.synchronize()
            step_ms2 = a0.elapsed_time(a1) / k
            scan_ms2 = s2.time_scans(k)
            other_paths[label] = {"ms_per_step": step_ms2, "value": P / (step_ms2 * 1e-3), "kernel_ms": scan_ms2,
                                  "kernel": "two_opt_scan_matrix_kernel", "scan_moves_per_s": P / (scan_ms2 * 1e-3)}
            s2.close()
            p2.close()

    # --- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e_moves = args.e2e_moves
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    hx = torch.from_numpy(x).pin_memory().numpy()
    hy = torch.from_numpy(y).pin_memory().numpy()
    htour = torch.from_numpy(start.astype(np.uint32).view(np.int32)).pin_memory().numpy().view(np.uint32)

    def e2e_step():
        p2 = T.Problem.euc2d(ctx, hx, hy)
        t2, st2, _ = p2.local_search(T.ALGO_TWO_OPT_BEST, htour, path=path, max_moves=e2e_moves)
        p2.close()
        return int(st2.evals), t2

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    evals = 0
    for _ in range(e2e_steps):
        ev, _t = e2e_step()
        evals += ev
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        tdt = torch.tensor([dt], device="cuda")
        dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        dt = float(tdt.item())
    e2e_value = world * evals / dt
    h2d = 2 * 4 * n + 4 * n  # x, y, start tour
    d2h = 4 * n + 64         # tour + stats

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # --- roofline for the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    traffic = {}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` captures
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass

    def hbm_roofline(kernel_ms, key):
        achieved = 4.0 * P / (kernel_ms * 1e-3) / 1e9  # 4 algorithmic bytes per move
        t = traffic.get(key, {}) if args.workload == "n10k" else {}
        return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": t.get("bytes_per_launch"), "traffic_source": t.get("source"), "peak_source": peak_src,
                "kernel": "two_opt_scan_matrix_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes_per_move": 4,
                "algorithmic_bytes_per_launch": 4 * P}

    for label, o in other_paths.items():
        o["roofline"] = hbm_roofline(o["kernel_ms"], label)
    if stats_path == T.PATH_MATRIX:
        roofline = hbm_roofline(scan_ms, "matrix_f32")
    else:
        ffma, mufu = ctx.microbench_fp32()
        flops = 15.0 * P / (scan_ms * 1e-3)  # 13 FP32 ops + 2 sqrt per move
        roofline = {"bound": "fp32_issue", "achieved": flops / 1e12, "peak": ffma / 1e12, "unit": "TFLOP/s",
                    "frac": flops / ffma,
                    "peak_source": "measured here: dependent-free FFMA stream, lane-instructions/s "
                                   "(tl_microbench_fp32); MEASURED_PEAKS.json has no FP32 figure",
                    "kernel": "two_opt_scan_recompute_kernel", "kernel_ms": scan_ms,
                    "algorithmic_flop_per_move": 15, "mufu_peak_per_s": mufu,
                    "traffic": traffic.get("recompute", {}).get("bytes_per_launch") if args.workload == "n10k" else None,
                    "traffic_source": traffic.get("recompute", {}).get("source"),
                    "note": "coordinate-recompute path: ~0 bytes/move, bound by FP32 issue, not HBM or tensor"}
    roofline["kernel_share_of_step"] = scan_ms / (ms / args.steps)

    # --- CPU baseline: the oracle port on the host cores (bounded sample)
    threads = os.cpu_count() or 1
    cpu_v, cpu_scans, cpu_dt = cpu_scan_rate(n, seed, threads, args.cpu_budget)
    cpu_v1, cpu_scans1, cpu_dt1 = cpu_scan_rate(n, seed, 1, min(args.cpu_budget, 4.0))
    cpu_baseline = {
        "value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{cpu_scans} full Mode B scans of the same n={n} tour in {cpu_dt:.1f} s, {threads} threads "
                  f"(row-parallel oracle scan)",
        "single_thread_value": cpu_v1,
        "note": "C oracle with flat arrays; the Rust reference is single-threaded and pays 2 SipHash "
                "lookups per distance, so this over-states the reference's speed",
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: n={n} uniform EUC_2D (splitmix64 seed {seed}), NN start tour, "
                               f"best-improvement 2-opt, one step = full scan of {P} moves + apply",
                   "path": "matrix" if stats_path == T.PATH_MATRIX else "recompute",
                   "l2": "inputs larger than L2 (400 MB matrix)" if stats_path == T.PATH_MATRIX else
                         "compute-bound kernel: 160 KB of tour-ordered points, L2 state irrelevant",
                   "parallelism": f"replicas x{world} (independent multi-start tours)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "step": f"tl_problem_create_euc2d + tl_local_search(max_moves={e2e_moves}) + tour read-back, "
                        f"{e2e_steps} calls, pinned host buffers", "seconds": dt},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "other_paths": other_paths,
        "moves_applied": real_steps,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="n10k", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="auto", choices=["auto", "recompute", "matrix"])
    ap.add_argument("--e2e-moves", type=int, default=100)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

"""Summarise an ncu report (run HERE, no GPU needed): `ncu_summary.py <report.ncu-rep> [out.md]`.
Writes the per-launch headline metrics, pipe utilisation and warp-stall breakdown as markdown; the
judge-facing copies live under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "L1 global-load hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy % (of active cycles)"),
    ("smsp__cycles_active.avg", "SMSP active cycles (avg)"), ("sm__cycles_elapsed.max", "SM elapsed cycles (max)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu summary of `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`)\n"]
    out.append("Per-launch values are cold-cache and serialised by the profiler (about 40 replays per launch): "
               "use them for traffic, instruction counts and stall shares, not for absolute time.\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out.append(f"## launch {r[hdr.index('ID')]}: `{name}`\n")
        out.append("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"| {label} (`{k}`) | {r[i]} {units[i]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") \
                    and "not_issued" not in h:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.05:
                    stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        out.append("\nWarp states per issued instruction (warps per scheduler, `smsp__average_warps_issue_stalled_*`): "
                   + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)) + "\n")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()

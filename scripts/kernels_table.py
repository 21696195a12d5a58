"""`kernels_table.py <ncu.csv> [out.md]`: per-kernel totals from an ncu --csv log of scripts/kernels_target.py."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    kid, name, metric, unit, val = r[0], r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
    short = name.split("(")[0].replace("void ", "").replace("tl::", "").replace("unnamed>::", "").replace("<unnamed>::", "")
    a = agg.setdefault(short, {"launches": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0, "inst": 0.0})
    a["launches"].add(kid)
    scale = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1}.get(unit, 1)
    if metric == "gpu__time_duration.sum":
        a["ns"] += val * scale
    elif metric == "dram__bytes_read.sum":
        a["rd"] += val * scale
    elif metric == "dram__bytes_write.sum":
        a["wr"] += val * scale
    elif metric == "smsp__inst_executed.sum":
        a["inst"] += val * scale
out = ["| kernel | launches | total time | DRAM read | DRAM write | DRAM GB/s | warp-instr | instr/clk/SM @1.965 GHz |", "|---|---|---|---|---|---|---|---|"]
for k, a in agg.items():
    t = a["ns"] * 1e-9
    out.append(f"| `{k}` | {len(a['launches'])} | {a['ns'] / 1e3:.1f} us | {a['rd'] / 1e6:.2f} MB | {a['wr'] / 1e6:.2f} MB | "
               f"{(a['rd'] + a['wr']) / t / 1e9 if t else 0:.0f} | {a['inst']:.3g} | {a['inst'] / (t * 1.965e9 * 148) if t else 0:.2f} |")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)

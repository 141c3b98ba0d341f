"""Summarise an `ncu --set full` report into markdown + traffic.json.
   ncu -i X.ncu-rep --page raw --csv > raw.csv ; python scripts/summarize_ncu.py raw.csv <ticks> profiles/NAME.md"""
import csv, json, os, sys

ALGO = {"k_dollar_tasks": 16, "k_dollar_chunk_sums": 16, "k_bar_ohlcv_median": 16, "k_bar_ohlcv_warp": 16, "k_bar_ohlcv_conveyor": 16, "k_bar_ohlcv_median_v1": 16, "k_bar_order_stats": 8}
rows = list(csv.reader(open(sys.argv[1])))
ticks = float(sys.argv[2])
out = sys.argv[3]
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
lines = [f"# ncu --set full summary ({os.path.basename(sys.argv[1])}, {ticks:.0e} ticks per launch, --clock-control none)\n"]
traffic = {}
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
def to_bytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * m.get(u, 1)
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    if name.startswith("void "): name = name[5:]
    shown = name
    name = {"k_dollar_tasks_t": "k_dollar_tasks"}.get(name.split("<")[0], name.split("<")[0] if name.split("<")[0] in ALGO else name)
    lines.append(f"## {shown}\n")
    lines.append("| metric | value |\n|---|---|")
    for key, label in want:
        if key in col:
            lines.append(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
    rd = to_bytes(num(r[col["dram__bytes_read.sum"]]), units[col["dram__bytes_read.sum"]])
    wr = to_bytes(num(r[col["dram__bytes_write.sum"]]), units[col["dram__bytes_write.sum"]])
    dur = num(r[col["gpu__time_duration.sum"]])
    du = units[col["gpu__time_duration.sum"]]
    dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(du, 1e-3)
    algo = ALGO.get(name)
    if algo:
        lines.append(f"| algorithmic bytes ({algo} B/tick) | {algo * ticks / 1e9:.3f} GB |")
        lines.append(f"| measured DRAM traffic / algorithmic | {(rd + wr) / (algo * ticks):.3f} |")
        lines.append(f"| algorithmic GB/s under ncu (cold, serialised) | {algo * ticks / dur_s / 1e9:.0f} |")
        traffic.setdefault(name, {})
        if not traffic[name]: traffic[name] = {"dram_bytes_per_launch": rd + wr, "ticks_per_launch": ticks, "bytes_per_tick": (rd + wr) / ticks}
    top = sorted(((num(r[col[s]]) or 0, s) for s in stalls), reverse=True)[:4]
    lines.append("| top stall reasons (warps per issue) | " + ", ".join(f"{s.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, s in top) + " |")
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
tj = os.path.join(os.path.dirname(out), "traffic_" + os.path.splitext(os.path.basename(out))[0] + ".json")
json.dump(traffic, open(tj, "w"), indent=1)
print("wrote", out, tj)

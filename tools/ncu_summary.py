"""Summarise an ncu --set full report (raw page CSV) into a markdown table: python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; data = rows[2:]
def col(name): return hdr.index(name)
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
units = rows[1]
agg = collections.OrderedDict()
ki = col("Kernel Name")
for r in data:
    name = r[ki].split("(")[0].replace("<unnamed>::", "")
    agg.setdefault(name, []).append(r)
print("| kernel | launches | " + " | ".join(n for _, n in want) + " |")
print("|---|---|" + "---|" * len(want))
for name, rs in agg.items():
    cells = []
    for m, _ in want:
        if m not in hdr:
            cells.append("-"); continue
        vals = [float(r[col(m)].replace(",", "")) for r in rs if r[col(m)] not in ("", "n/a")]
        if not vals:
            cells.append("-"); continue
        u = units[col(m)]
        v = sum(vals) if m.endswith(".sum") and "time" in m else (sum(vals) if "bytes" in m else sum(vals) / len(vals))
        cells.append(f"{v:.3g} {u}".strip())
    print(f"| `{name}` | {len(rs)} | " + " | ".join(cells) + " |")

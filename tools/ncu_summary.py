"""Per-kernel summary table from `ncu -i X.ncu-rep --page raw --csv` (a --set full capture)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %")]
cols = [(c, n) for c, n in cols if c in hdr]
print(f"# {sys.argv[1]}: ncu --set full --clock-control none (cold-cache, serialised)")
print(f"{'kernel':44s} " + " ".join(f"{n:>12s}" for _, n in cols))
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:44]
    vals = []
    for c, _ in cols:
        i = hdr.index(c)
        v = r[i]
        try:
            f = float(v)
            v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.3e}"
        except ValueError:
            pass
        vals.append(f"{v + ' ' + units[i][:5]:>12s}")
    print(f"{name:44s} " + " ".join(vals))

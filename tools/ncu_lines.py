"""Per-source-line instruction share from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines, cur, ci, sa = [], "", None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ci, sa = r.index("Instructions Executed"), r.index("# Samples")
    elif ci is not None and r[0].isdigit() and len(r) > ci:
        try:
            lines.append((cur, r[0], r[1], float(r[ci] or 0), float(r[sa] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print(f"total warp instructions {tot:.0f}, samples {ts:.0f}")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{l[3] / tot * 100:5.1f}% inst {l[4] / max(ts,1) * 100:5.1f}% samp  {l[0]}:{l[1]:>4}  {l[2][:110]}")

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel
table (count, total, average, share). Per-launch times under ncu are cold-cache and
serialised: compare SHARES, not absolutes."""
import collections
import csv
import io
import sys


def main(path, out=None):
    txt = open(path).read()
    start = txt.index('"ID"')
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(txt[start:])):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1e-3)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    lines = [f"# source: {path}", f"# total kernel time {tot:.1f} us over {sum(a[0] for a in agg.values())} launches",
             f"{'kernel':72s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'share':>7s}"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:72s} {n:5d} {t:11.1f} {t / n:9.2f} {100 * t / tot:6.1f}%")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

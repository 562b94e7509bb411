"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel → markdown."""
import collections
import csv
import sys


def main(path, out):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "")
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = row.get("Kernel Name", "").split("(")[0][-90:]
        agg[name][0] += 1
        agg[name][1] += v
        total += v
    with open(out, "w") as f:
        f.write(f"# Per-kernel device time (ncu gpu__time_duration.sum, serialised launches; compare SHARES)\n\n")
        f.write(f"source: `{path}`, total {total / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches\n\n")
        f.write("| time (us) | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
            f.write(f"| {t:.1f} | {100 * t / total:.1f}% | {n} | `{k}` |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (name, grid)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
tokens = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void |b200::|\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", name)
    a = agg.setdefault((name, row["Grid Size"], row["Block Size"]), [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us of kernel time "
      f"({tot / tokens:.1f} us per token over {tokens:g} tokens; cold-cache, serialised under ncu)")
print("| kernel | grid | block | launches | avg us | total us | share |")
print("|---|---|---|---|---|---|---|")
for (name, grid, block), (n, t) in agg.items():
    print(f"| `{name}` | {grid} | {block} | {n} | {t / n:.2f} | {t:.1f} | {100 * t / tot:.1f}% |")

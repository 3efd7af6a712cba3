"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): per kernel (name, grid, block) the launch
count, mean/total device time and share of the total. usage: python scripts/launch_summary.py launches.csv [tokens]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 1
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    if "<" in r[4]:
        name = r[4].split("(b200")[0].split("(Matvec")[0].split("(Attn")[0]
    key = (name, r[8], r[7])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14])
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tokens} token(s), total device time {tot / 1e3:.1f} us  ({tot / 1e3 / tokens:.1f} us/token)")
print(f"{'kernel':44s} {'grid':>14s} {'block':>13s} {'n':>5s} {'mean us':>9s} {'total us':>10s} {'share':>7s}")
for (name, grid, block), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:44s} {grid:>14s} {block:>13s} {n:5d} {t / n / 1e3:9.2f} {t / 1e3:10.1f} {100 * t / tot:6.1f}%")

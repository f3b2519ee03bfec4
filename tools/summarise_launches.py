"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (template arguments kept), block and grid:
python tools/summarise_launches.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
acc = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("opalb200::", "").replace("(opalb200::Lanes)", "")
    name = re.sub(r"void |at::native::|<unnamed>::", "", name)
    key = (name[:70], r[7], r[8])
    acc[key][0] += 1
    acc[key][1] += float(r[14]) / 1e3
total = sum(v[1] for v in acc.values())
print(f"{len(rows)} launches, {total / 1e3:.1f} ms of kernel time (serialised by ncu)")
print(f"{'kernel':70s} {'block':>14s} {'grid':>16s} {'n':>5s} {'sum ms':>9s} {'avg us':>9s} {'share':>6s}")
for (name, blk, grd), (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:70s} {blk:>14s} {grd:>16s} {n:5d} {us / 1e3:9.2f} {us / n:9.1f} {100 * us / total:5.1f}%")

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and share.
usage: python tools/summarize_launches.py launches.csv [launches_per_step]  (keeps the LAST launches_per_step rows if given)"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
recs = list(csv.DictReader(rows))
if len(sys.argv) > 2:
    recs = recs[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for r in recs:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
print("launches %d, sum of kernel durations %.1f us (ncu: cold caches, serialised)" % (len(recs), tot))
print("%-64s %6s %11s %7s" % ("kernel", "count", "total us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-64s %6d %11.1f %6.1f%%" % (k[:64], v[0], v[1], 100 * v[1] / tot))

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg, tot = None, collections.OrderedDict(), 0.0
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", d["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:80]
        v = float(d["Metric Value"].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
print(f"total {tot / 1e3:.3f} ms in {sum(a[0] for a in agg.values())} launches")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}% {n:5d}  {k}")

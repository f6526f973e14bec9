"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name: launches, total us, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[h + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = r[ki].split("(")[0].split("::")[-1]
    if name.startswith("pack_"):
        continue
    agg[name][0] += v
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print("total %.2f ms, %d launches" % (tot / 1000, sum(v[1] for v in agg.values())))
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%6.1f %%  %9.1f us  %5d x  %s" % (100 * v / tot, v, c, n))

"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of profiles/profile_step.py:
per pass (split at the temb_kernel that starts every forward program and the first VJP kernel)
the time share of every kernel family."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = []
for r in rows[h + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = r[ki].split("(")[0].split("::")[-1]
    data.append((name, v))
# split into programs: a forward program starts with temb_kernel; the VJP program with the first
# kernel after the forward's final edge_reduce
names = sys.argv[2:] or ["jvp", "vjp", "fwd_b1", "fwd_b8"]
passes, cur = [], []
for n, v in data:
    if n.startswith("pack_") or n.startswith("set_scalar"):
        continue
    # (temb_kernel until the last session of round 2, then two temb_dense_kernel launches: split at the first)
    if (n.startswith("temb_kernel") or (n.startswith("temb_dense_kernel") and not (cur and cur[-1][0].startswith("temb_dense_kernel")))) and cur:
        passes.append(cur); cur = []
    cur.append((n, v))
    if n.startswith("edge_reduce") and len(cur) > 50 and any(x[0].startswith(("temb_kernel", "temb_dense_kernel")) for x in cur):
        passes.append(cur); cur = []
if cur:
    passes.append(cur)
for i, p in enumerate(passes):
    tot = sum(v for _, v in p)
    agg = collections.defaultdict(lambda: [0.0, 0])
    for n, v in p:
        agg[n][0] += v; agg[n][1] += 1
    print("== %s: %.2f ms, %d launches" % (names[i] if i < len(names) else "pass%d" % i, tot / 1000, len(p)))
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("  %9.2f ms %5.1f%% n=%4d avg=%7.1fus %s" % (t / 1000, 100 * t / tot, c, t / c, n))

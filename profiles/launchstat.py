"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals.
usage: python profiles/launchstat.py launches.csv [skip_first_n] [take_n]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    data.append((int(r[ix["ID"]]), r[ix["Kernel Name"]], us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
take = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
data = data[skip:skip + take]
agg = OrderedDict()
for _, k, us in data:
    k = k[:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("launches %d, total %.1f us" % (len(data), tot))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("  %6.1f us  %5.1f%%  x%-3d %s" % (us, 100 * us / tot, n, k))

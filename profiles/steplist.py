"""Ordered kernel list of ONE replayed SVI step from an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv --log-file L.csv python bench.py ...): the launches between two
consecutive GLM kernels.  usage: python profiles/steplist.py L.csv [anchor-substring] [which]"""
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [(r[ix["Kernel Name"]], float(r[ix["Metric Value"]].replace(",", "")) / 1e3, r[ix["Grid Size"]])
        for r in rows[hi + 1:] if len(r) >= len(hdr)]
anchor = sys.argv[2] if len(sys.argv) > 2 else "glm_bernoulli_mma"
idx = [i for i, d in enumerate(data) if anchor in d[0]]
which = int(sys.argv[3]) if len(sys.argv) > 3 else min(10, len(idx) - 2)
a, b = idx[which], idx[which + 1]
tot = 0.0
for k, us, g in data[a:b]:
    k = re.sub(r"void |at::native::|at::|<unnamed>::", "", k)[:110]
    tot += us
    print("%7.1f us  grid %-14s %s" % (us, g, k))
print("launches %d, total %.1f us (ncu per-launch times: cold caches, serialised)" % (b - a, tot))

"""Summarise an `ncu --page source --csv` dump: opcode mix weighted by executions + stall reasons.
usage: ncu -i X.ncu-rep --page source --csv > x.csv; python profiles/srcstat.py x.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
src, ex, ws = ix["Source"], ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
c, cs = Counter(), Counter()
stalls = Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        n = int(r[ex]); s = int(r[ws] or 0)
    except ValueError:
        continue
    parts = r[src].split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = op.split(".")[0]
    c[op] += n; cs[op] += s; tot += n
    for h in stall_cols:
        try:
            stalls[h] += int(r[ix[h]] or 0)
        except ValueError:
            pass
print("kernel:", rows[0][1] if rows[0] else "?")
print("total warp instructions executed:", tot)
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print("  %-10s %12d %5.1f%%   stall samples %d" % (op, n, 100.0 * n / max(tot, 1), cs[op]))
st = sum(stalls.values())
print("stall reasons (samples):", ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(st, 1)) for k, v in stalls.most_common(8)))

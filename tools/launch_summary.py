"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ms, share."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else (v * 1e3 if u in ("s", "second") else v))
    name = r[ki].split("(")[0][:60]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("%-62s %6s %12s %7s" % ("kernel", "count", "total ms", "share"))
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-62s %6d %12.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("%-62s %6d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))

"""Summarise an ncu source-page CSV: stall samples by reason, and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def f(x):
    try: return float(x)
    except: return 0.0
tot = {h: sum(f(r[col[h]]) for r in data) for h in stall_cols}
alls = sum(tot.values())
print("total samples", alls)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-26s %8.0f  %5.1f%%" % (h, v, 100 * v / alls))
top = sorted(data, key=lambda r: -f(r[col["# Samples"]]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
print("hottest instructions (samples, executed, top stall, sass):")
for r in top:
    st = max(stall_cols, key=lambda h: f(r[col[h]]))
    print("  %6.0f %10s %-18s %s" % (f(r[col["# Samples"]]), r[col["Instructions Executed"]], st, r[col["Source"]][:90]))

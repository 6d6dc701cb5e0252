"""Top SASS instructions of an ncu source-page CSV by executed count / samples; and opcode histogram weighted by executions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iconf, iwf, iideal = hdr.index("L1 Conflicts Shared N-Way"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
recs = []
for r in rows[2:]:
    try:
        recs.append((int(r[iex] or 0), int(r[isamp] or 0), r[isrc].strip(), r[iconf], r[iwf], r[iideal], r[ia]))
    except Exception:
        pass
tot = sum(x[0] for x in recs); totS = sum(x[1] for x in recs)
print("total warp-instr executed", tot, "samples", totS)
hist = collections.Counter(); hs = collections.Counter()
for ex, sm, src, *_ in recs:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    hist[op] += ex; hs[op] += sm
print("opcode: executed share | sample share")
for op, c in hist.most_common(22):
    print(f"  {op:10s} {100*c/tot:6.2f}%  {100*hs[op]/max(totS,1):6.2f}%")
mode = sys.argv[2] if len(sys.argv) > 2 else "samples"
key = (lambda x: -x[1]) if mode == "samples" else (lambda x: -x[0])
print("top instructions by", mode)
for ex, sm, src, conf, wf, ideal, addr in sorted(recs, key=key)[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"  ex={ex:11d} samp={sm:6d} conf={conf:>6s} wf={wf:>10s} ideal={ideal:>10s}  {src[:90]}")

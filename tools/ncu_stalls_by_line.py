"""Per-source-line stall-sample table of an ncu report captured with --import-source on (lineinfo build).
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_stalls_by_line.py src.csv [kernel-substr] [topN]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur, hdr, fn, seen_fn = None, None, None, []
cols = ["# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_branch_resolving", "stall_math", "stall_barrier"]
agg = {c: collections.Counter() for c in cols}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        if fn not in seen_fn:
            seen_fn.append(fn)
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or want not in (fn or ""):
        continue
    if len(seen_fn) > 1 and fn != [f for f in seen_fn if want in f][0]:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (cur, ln, r[1].strip()[:100])
    for c in cols:
        v = r[ix[c]]
        agg[c][key] += int(v) if v not in ("", "-") else 0
S = agg["# Samples"]
tot = sum(S.values())
print("kernels in report:", seen_fn)
print("total samples", tot, "warp instructions", sum(agg["Instructions Executed"].values()))
byfile = collections.Counter()
for k, v in S.items():
    byfile[k[0]] += v
print("by file:", byfile.most_common())
print("--- top lines by samples:  samples  %  inst | long_sb short_sb wait no_inst branch math barrier")
for k, v in S.most_common(topn):
    print(f"{v:6d} {100 * v / max(tot, 1):5.1f}% {agg['Instructions Executed'][k]:9d} | " +
          " ".join(f"{agg[c][k]:5d}" for c in cols[2:]) + f"  {k[0]}:{k[1]}  {k[2]}")
for c in ("stall_long_sb", "stall_no_inst"):
    print(f"--- top lines by {c}")
    for k, v in agg[c].most_common(20):
        print(f"{v:6d} inst {agg['Instructions Executed'][k]:8d}  {k[0]}:{k[1]}  {k[2]}")

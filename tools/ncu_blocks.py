"""Condensed view of an ncu source-page CSV: basic blocks (runs of SASS lines with the
same execution count) with their share of executed instructions and of stall samples.
usage: ncu_blocks.py <source.csv> [kernel index] [min share %]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; ks.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
k = ks[ki]; h = k["hdr"]; iS = h.index("Source"); iE = h.index("Instructions Executed"); iSm = h.index("# Samples")
tot = sum(int(r[iE]) for r in k["rows"]); totS = sum(int(r[iSm]) for r in k["rows"])
print(k["name"], "total inst", tot, "nlines", len(k["rows"]), "samples", totS)
blocks = []
for idx, r in enumerate(k["rows"]):
    e = int(r[iE]); s = int(r[iSm])
    if blocks and blocks[-1]["e"] == e: b = blocks[-1]; b["n"] += 1; b["s"] += s; b["end"] = idx
    else: blocks.append({"e": e, "n": 1, "s": s, "start": idx, "end": idx})
for b in blocks:
    if 100 * b["e"] * b["n"] / tot < thr and 100 * b["s"] / totS < 1.0: continue
    ops = {}
    for r in k["rows"][b["start"]:b["end"] + 1]:
        t = r[iS].split(); op = t[1] if t[0].startswith('@') else t[0]
        op = op.split('.')[0]; ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:7]
    print(f"lines {b['start']:5d}-{b['end']:5d} n={b['n']:4d} exec={b['e']:10d} inst%={100*b['e']*b['n']/tot:5.1f} samp%={100*b['s']/totS:5.1f}  {top}")

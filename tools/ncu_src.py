#!/usr/bin/env python3
"""tools/ncu_src.py <report> <kernel-regex> [min_pct]: SASS of one launch cut into runs of equal execution count."""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iS, iE, iN, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
data = [(r[iS].strip(), int(r[iE]), int(r[iN]), int(r[iT])) for r in rows[h + 1:] if len(r) > iT and r[iE].isdigit()]
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print("instructions %d  warp-inst %d  thread-inst %d  samples %d" % (len(data), tot, sum(d[3] for d in data), ts))
start = 0
def flush(a, b):
    seg = data[a:b]
    e = sum(d[1] for d in seg); s = sum(d[2] for d in seg)
    if 100.0 * e / tot >= minpct or 100.0 * s / max(ts, 1) >= minpct:
        ops = {}
        for d in seg:
            op = d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
        top = ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
        print("  [%4d,%4d) n=%4d exec/instr %9d  inst %5.1f%%  samples %5.1f%%  lanes %.1f | %s" % (
            a, b, b - a, seg[0][1], 100.0 * e / tot, 100.0 * s / max(ts, 1), sum(d[3] for d in seg) / max(e, 1), top))
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][1] - data[start][1]) > 0.25 * max(data[start][1], 1):
        flush(start, i); start = i

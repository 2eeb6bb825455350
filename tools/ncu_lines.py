#!/usr/bin/env python3
"""tools/ncu_lines.py <report> <kernel-regex> [top]: stall samples and instructions per CUDA source line."""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No" and "# Samples" in r)
hdr = rows[h]
iN, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
agg = collections.OrderedDict()
cur = None
for r in rows[h + 1:]:
    if len(r) <= iN: continue
    if r[0].strip():
        cur = (r[0], r[1].strip())
    if cur is None or not r[iN].isdigit(): continue
    a = agg.setdefault(cur, [0, 0]); a[0] += int(r[iN]); a[1] += int(r[iE]) if r[iE].isdigit() else 0
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
print("samples %d warp-inst %d" % (ts, te))
for (ln, src), (s_, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% inst  L%-5s %s" % (100.0 * s_ / max(ts, 1), 100.0 * e / max(te, 1), ln, src[:105]))

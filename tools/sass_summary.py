#!/usr/bin/env python3
"""tools/sass_summary.py [library.so] > profiles/r2_sass_summary.txt
Per kernel of the product library: SASS instruction count and a histogram of mnemonics (cuobjdump -sass), with the
ones that prove the asynchronous copy paths called out (UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier,
LDGSTS = cp.async)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "aws-c-compression_b200", "lib", "libaws-c-compression.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("SASS summary of %s (sm_100a)\n" % os.path.relpath(lib, ROOT))
flags = ("UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "BAR", "ATOMS", "ATOMG", "RED")
for k, h in hist.items():
    if not k:
        continue
    total = sum(h.values())
    top = ", ".join("%s %d" % kv for kv in h.most_common(10))
    mark = ", ".join("%s x%d" % (f, h[f]) for f in flags if h[f])
    print("%-52s %5d instr | %s" % (k[:52], total, top))
    if mark:
        print("%-52s       | async / sync: %s" % ("", mark))

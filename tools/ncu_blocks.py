#!/usr/bin/env python
"""Executed-instruction blocks of one kernel from an ncu report: runs of SASS instructions with the same
execution count, with their opcode histogram.  Shows which rare paths are entered how often.

    python tools/ncu_blocks.py <report.ncu-rep> [min_share_percent]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, data = rows[1], rows[2:]
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
total = sum(int(r[i_ex] or 0) for r in data)
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
blocks = []
prev, start, ops = None, 0, []
for i, r in enumerate(data + [None]):
    e = int(r[i_ex] or 0) if r else None
    if e != prev:
        if prev is not None:
            blocks.append((start, i, prev, ops))
        prev, start, ops = e, i, []
    if r:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src].strip())
        ops.append(m.group(2) if m else "?")
loop = max(b[2] for b in blocks if len(b[3]) > 50)
print("total warp-instructions %d; hot-loop iterations %d; per iteration %.1f" % (total, loop, total / loop))
f64 = sum(int(r[i_ex] or 0) for r in data if re.match(r"(@!?U?P\d+\s+)?(DFMA|DMUL|DADD|DSETP|DMNMX)", r[i_src].strip()))
print("FP64 per iteration %.1f, other %.1f, 2F+O = %.1f" % (f64 / loop, (total - f64) / loop, (total + f64) / loop))
for s, e, cnt, ops in blocks:
    share = 100.0 * cnt * len(ops) / total
    if share < min_share:
        continue
    c = collections.Counter(ops)
    nf = sum(v for k, v in c.items() if k in FP64)
    print("%6s..%6s x%9d (%5.1f%% of iters) n=%3d fp64=%3d  %5.2f%%  %s" % (
        data[s][0][-5:], data[e - 1][0][-5:], cnt, 100.0 * cnt / loop, len(ops), nf, share,
        " ".join("%s:%d" % kv for kv in c.most_common(8))))

#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction execution counts to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <object-or-cubin> <kernel-name-substring> [warp_iterations]

Joins `ncu --page source --csv` (addresses + "Instructions Executed") with `nvdisasm -gi` line info of
the same kernel, by instruction offset, and prints executed warp-instructions per source line split
into FP64 / other.  Used to find where the fused kernel spends its issue slots (profiles/*.md)."""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, obj, kname = sys.argv[1:4]
    witer = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = rows[2:]
    base = int(data[0][ia], 16)
    counts = {int(r[ia], 16) - base: (r[isrc].strip(), int(r[iex])) for r in data}

    tmp = tempfile.mkdtemp()
    if obj.endswith(".cubin"):
        cubin = obj
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
        cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    # find the function section
    sec = None
    cur_line = "?"
    per_line = collections.defaultdict(lambda: [0, 0])
    infn = False
    for line in dis.splitlines():
        if line.startswith("//--------------------- .text."):
            infn = kname in line
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            if "inlined at" in m.group(3):
                # keep the outermost location in erk_fused.cuh when present
                inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                locs = [(m.group(1), m.group(2))] + inl
            else:
                locs = [(m.group(1), m.group(2))]
            pick = None
            for f, l in locs[::-1]:
                if f.endswith("erk_fused.cuh"):
                    pick = (f, l)
                    break
            if pick is None:
                pick = locs[-1]
            cur_line = "%s:%s" % (os.path.basename(pick[0]), pick[1])
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m:
            off = int(m.group(1), 16)
            if off in counts:
                src, ex = counts[off]
                op = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', src)
                op = op.group(2) if op else "?"
                fp64 = op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
                per_line[cur_line][0 if fp64 else 1] += ex
    tot = [sum(v[0] for v in per_line.values()), sum(v[1] for v in per_line.values())]
    print("total fp64 %.1f other %.1f (per warp-iteration, %g iterations)" % (tot[0] / witer, tot[1] / witer, witer))
    for k, v in sorted(per_line.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:60]:
        print("%-28s fp64 %8.1f  other %8.1f" % (k, v[0] / witer, v[1] / witer))


if __name__ == "__main__":
    main()

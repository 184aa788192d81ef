#!/usr/bin/env python
"""Opcode mix and stall totals per kernel from an .ncu-rep source page:
python scripts/ncu_sass_mix.py file.ncu-rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# the page is a sequence of blocks: ["Kernel Name", name], header row, instruction rows
blocks = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        blocks.append([rows[i][1], rows[i + 1], []])
        i += 2
        continue
    if blocks and len(rows[i]) >= len(blocks[-1][1]):
        blocks[-1][2].append(rows[i])
    i += 1
for name, hdr, body in blocks:
    ix = {k: j for j, k in enumerate(hdr)}
    ops = collections.Counter(); stall = collections.Counter()
    tot = 0
    for r in body:
        toks = r[ix["Source"]].split()
        if not toks: continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        base = op.split('.')[0]
        if base.startswith(('LD', 'ST', 'BAR', 'ATOM')): base = '.'.join(op.split('.')[:2])
        n = int(r[ix["Instructions Executed"]] or 0)
        ops[base] += n; tot += n
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                stall[k] += int(r[ix[k]] or 0)
    print("==", name, "--", len(body), "SASS instructions,", tot, "warp instructions executed")
    print("  " + ", ".join("%s %.1f%%" % (op, 100.0 * n / max(tot, 1)) for op, n in ops.most_common(14)))
    ts = max(sum(stall.values()), 1)
    print("  stalls: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * n / ts) for k, n in stall.most_common(7)))

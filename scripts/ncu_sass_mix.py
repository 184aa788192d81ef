#!/usr/bin/env python
"""Opcode mix and stall totals from an .ncu-rep source page: python scripts/ncu_sass_mix.py file.ncu-rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); stall = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0] + ('.' + '.'.join(op.split('.')[1:2]) if op.startswith(('LD', 'ST', 'BAR', 'ATOM')) else '')
    n = int(r[ix["Instructions Executed"]] or 0)
    ops[op] += n; tot += n
    samp[op] += int(r[ix["# Samples"]] or 0)
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            stall[k] += int(r[ix[k]] or 0)
print("total warp insts", tot)
for op, n in ops.most_common(28):
    print("  %-14s %12d %5.1f%%   samples %6d" % (op, n, 100.0 * n / tot, samp[op]))
ts = sum(stall.values())
print("stalls (all samples):")
for k, n in stall.most_common(10):
    print("  %-26s %5.1f%%" % (k, 100.0 * n / ts))

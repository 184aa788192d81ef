#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep captured with --import-source on:
   python scripts/ncu_src_hot.py file.ncu-rep <kernel regex> [top N] [launch index]
Prints the SASS instructions with the most stall samples and their dominant stall reason."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
# several kernels are concatenated: split at "Kernel Name" lines
blocks, cur = [], []
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur:
            blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
if cur:
    blocks.append(cur)
blk = blocks[which]
print(blk[0])
rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[1:]:
    try:
        n = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot += n
    data.append((n, r))
bysum = {}
for n, r in data:
    for h in stall_cols:
        try:
            bysum[h] = bysum.get(h, 0) + int(r[ix[h]])
        except ValueError:
            pass
print("total samples", tot, {k: v for k, v in sorted(bysum.items(), key=lambda kv: -kv[1])[:8]})
# cumulative position: where in the instruction stream are the samples
acc = 0
marks = []
for i, (n, r) in enumerate(data):
    acc += n
    marks.append(acc)
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[ix[h]]) if r[ix[h]].isdigit() else 0, h) for h in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s" % (n, 100.0 * n / tot, r[ix["Source"]][:70], ", ".join("%s=%d" % (h[6:], v) for v, h in st)))

#!/usr/bin/env python
"""Stall samples per PHASE of a kernel: the SASS stream of one launch (ncu --import-source on) is cut at
barriers (BAR.SYNC / WARPSYNC optional) and every segment reports its samples, dominant stall reasons
and instruction mix.   python scripts/ncu_src_phases.py file.ncu-rep <kernel regex> [launch index]"""
import csv, io, subprocess, sys, re
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
blocks, cur = [], []
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur:
            blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
blocks.append(cur)
blk = blocks[which]
print(blk[0])
rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs = []
cur = dict(n=0, st={}, mix={}, first=None, exec=0)
def close():
    global cur
    if cur["first"] is not None:
        segs.append(cur)
    cur = dict(n=0, st={}, mix={}, first=None, exec=0)
tot = 0
for r in rows[1:]:
    try:
        n = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    sass = r[ix["Source"]].strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", sass).split()[0] if sass else "?"
    base = op.split(".")[0]
    if cur["first"] is None:
        cur["first"] = r[ix["Address"]]
    cur["n"] += n
    tot += n
    try:
        cur["exec"] += int(r[ix["Instructions Executed"]])
    except ValueError:
        pass
    for h in stall_cols:
        try:
            cur["st"][h[6:]] = cur["st"].get(h[6:], 0) + int(r[ix[h]])
        except ValueError:
            pass
    key = base if base in ("LDG", "STG", "LDS", "STS", "DFMA", "DMUL", "DADD", "BAR", "LDL", "STL", "ATOMG", "RED", "SHFL", "MUFU", "LDGSTS", "UTMALDG", "SYNCS") else None
    if key:
        cur["mix"][key] = cur["mix"].get(key, 0) + 1
    if base == "BAR":
        close()
close()
print("total samples", tot)
for i, s in enumerate(segs):
    st = sorted(s["st"].items(), key=lambda kv: -kv[1])[:3]
    print("seg %2d %6d %5.1f%% exec=%9d  %-46s %s" % (i, s["n"], 100.0 * s["n"] / max(tot, 1), s["exec"],
          " ".join("%s=%d" % kv for kv in st), " ".join("%s:%d" % kv for kv in sorted(s["mix"].items()))))

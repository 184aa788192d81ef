#!/usr/bin/env python
"""Per SOURCE LINE totals of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build):
   python scripts/ncu_src_lines.py file.ncu-rep <kernel regex> [top N] [launch index]
samples, executed warp instructions, shared-memory wavefronts (ideal / excessive) and local-memory accesses."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# split into launches at "Function Name" changes is not needed for a single launch; files are separated by "File Path"
cur_file, hdr, ix = None, None, None
agg = {}
launch = -1
seen_files = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def g(name):
        try:
            return int(r[ix[name]])
        except (KeyError, ValueError, IndexError):
            return 0
    k = (cur_file, line)
    a = agg.setdefault(k, dict(src=r[1].strip(), smp=0, ins=0, wf=0, wfx=0, loc=0))
    a["smp"] += g("# Samples")
    a["ins"] += g("Instructions Executed")
    a["wf"] += g("L1 Wavefronts Shared")
    a["wfx"] += g("L1 Wavefronts Shared Excessive")
    if r[ix["Address Space"]].startswith("Local"):
        a["loc"] += g("Instructions Executed")
tot = sum(a["smp"] for a in agg.values()) or 1
toti = sum(a["ins"] for a in agg.values()) or 1
print("total samples", tot, "warp instructions", toti,
      "shared wavefronts", sum(a["wf"] for a in agg.values()), "excessive", sum(a["wfx"] for a in agg.values()))
print("%-22s %6s %6s %6s %10s %10s  %s" % ("file:line", "smp%", "ins%", "local", "smem wf", "excess", "source"))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    print("%-22s %6.2f %6.2f %6s %10d %10d  %s" % (f + ":" + str(l), 100.0 * a["smp"] / tot, 100.0 * a["ins"] / toti,
                                                 "L" if a["loc"] else "", a["wf"], a["wfx"], a["src"][:90]))

#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]`
log per kernel name:  python scripts/ncu_launch_summary.py launches.csv [signals_per_launch] [--json out.json]"""
import collections, csv, json, sys
path = sys.argv[1]
nsig = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 0
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; ix = {k: i for i, k in enumerate(hdr)}
T = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
Bs = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
acc = collections.OrderedDict()
first_only = '--first' in sys.argv  # use only the first launch of every kernel name
seen_id = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    k = r[ix['Kernel Name']].split('(')[0].replace('void ', '')[:48]
    m = r[ix['Metric Name']]; u = r[ix['Metric Unit']]; v = float(r[ix['Metric Value']].replace(',', ''))
    if first_only and seen_id.setdefault(k, r[ix['ID']]) != r[ix['ID']]: continue
    a = acc.setdefault(k, {'n': 0, 'ms': 0.0, 'rd': 0.0, 'wr': 0.0})
    if m == 'gpu__time_duration.sum': a['n'] += 1; a['ms'] += v * T.get(u, 1e-3)
    elif m == 'dram__bytes_read.sum': a['rd'] += v * Bs.get(u, 1)
    elif m == 'dram__bytes_write.sum': a['wr'] += v * Bs.get(u, 1)
tot = sum(a['ms'] for a in acc.values())
print("| kernel | launches | total ms | share | DRAM read MB/launch | DRAM write MB/launch |" + (" DRAM bytes/signal |" if nsig else ""))
print("|---|---|---|---|---|---|" + ("---|" if nsig else ""))
out = {}
for k, a in acc.items():
    line = "| `%s` | %d | %.3f | %.1f %% | %.1f | %.1f |" % (k, a['n'], a['ms'], 100 * a['ms'] / tot, a['rd'] / a['n'] / 1e6, a['wr'] / a['n'] / 1e6)
    if nsig:
        line += " %.0f |" % ((a['rd'] + a['wr']) / a['n'] / nsig)
        out[k] = {"launches": a['n'], "ms_per_launch": a['ms'] / a['n'], "dram_bytes_per_signal": (a['rd'] + a['wr']) / a['n'] / nsig}
    print(line)
print("\ntotal %.3f ms" % tot)
if '--json' in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)

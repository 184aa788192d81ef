#!/usr/bin/env python
"""profiles/tree_dram_bytes.json from the ncu CSV written by scripts/ncu_traffic.sh:
   python scripts/make_tree_dram_json.py gpurun_out/<tag>_traffic.csv 1024 > profiles/tree_dram_bytes.json
Every launch of the LAST full-size step is listed with its measured dram__bytes_read.sum + dram__bytes_write.sum
(no modelled entries); bench.py multiplies tree_bytes_per_signal by the signals of a step for roofline.traffic."""
import csv, json, sys
path, batch = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    key = (r[ix["ID"]], r[ix["Kernel Name"]], r[ix["Grid Size"]])
    d = launches.setdefault(key, {})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)
    d[r[ix["Metric Name"]]] = val * scale
def grid(g):
    return int(g.strip("()").split(",")[0])
# the last full-size step: from the last k_tree_low2 launch with the largest grid to the end of the capture
seq = sorted(((int(i), name, grid(g), d) for (i, name, g), d in launches.items()))
gmax = max(g for _, n, g, _ in seq if "k_tree_low2" in n)
start = max(k for k, (_, n, g, _) in enumerate(seq) if "k_tree_low2" in n and g == gmax)
out, tree_total, cz_total = [], 0.0, 0.0
for i, name, g, d in seq[start:]:
    if "k_tree_low2" in name and g != gmax:
        break  # a smaller call (parity check) follows
    b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    short = name.replace("void ", "").split("(")[0]
    out.append({"kernel": short, "grid": g, "ms": round(d.get("gpu__time_duration.sum", 0), 4),
                "dram_read_bytes_per_signal": round(d.get("dram__bytes_read.sum", 0) / batch),
                "dram_write_bytes_per_signal": round(d.get("dram__bytes_write.sum", 0) / batch),
                "l2_bytes_per_signal": round(d.get("lts__t_bytes.sum", 0) / batch)})
    if "k_cz2" in name:
        cz_total += b / batch
    else:
        tree_total += b / batch
print(json.dumps({"tree_bytes_per_signal": round(tree_total), "chirpz_bytes_per_signal": round(cz_total),
                  "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of every launch of one config-2 step at %d signals per "
                            "launch, current kernels (scripts/ncu_traffic.sh, scripts/make_tree_dram_json.py); first-row-only "
                            "spectrum-carry path, all entries measured" % batch,
                  "launches": out}, indent=1))

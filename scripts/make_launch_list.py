#!/usr/bin/env python
"""profiles/r02_launches_batch1024.md from profiles/tree_dram_bytes.json (ncu, scripts/ncu_traffic.sh) and a bench JSON
line (CUDA-event times):  python scripts/make_launch_list.py profiles/r02_bench.json > profiles/r02_launches_batch1024.md"""
import json, sys
d = json.load(open("profiles/tree_dram_bytes.json"))
b = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
ev = b["roofline"]["kernel_ms_per_step"]
tot_ev = sum(ev.values())
def name(e):
    k = e["kernel"]
    if e["grid"] < 100:
        return "cz_filter"
    for pat, n in (("k_tree_low2", "tree_low2"), ("k_up_smem<11", "tree_up_smem_N2048"), ("k_up_smem<12", "tree_up_smem_N4096"),
                   ("k_up_smem<13", "tree_up_smem_N8192"),
                   ("k_up_smem_cluster<13", "tree_up_smem_N8192"), ("k_up_smem_cluster<14", "tree_up_smem_N16384"), ("k_up_rows_a", "tree_up_rows_a"), ("k_up_cols_cz", "tree_up_cols_cz"),
                   ("k_up_cols", "tree_up_cols"), ("k_up_rows_c", "tree_up_rows_c"), ("k_cz2_cols_fwd", "cz_cols_fwd"),
                   ("k_cz2_rows", "cz_rows"), ("k_cz2_cols_inv", "cz_cols_inv")):
        if k.startswith(pat):
            return n
    return k
agg = {}
for e in d["launches"]:
    a = agg.setdefault(name(e), [0, 0.0, 0])
    a[0] += 1
    a[1] += e["ms"]
    a[2] += e["dram_read_bytes_per_signal"] + e["dram_write_bytes_per_signal"]
tot = sum(a[1] for a in agg.values())
print("# Round 2 -- ncu launch list of one config-2 step (current kernels: L2 prefetch, cluster kernels for N = 8192 / 16384, fused last column pass)\n")
print("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none`")
print("around `python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline --no-extras` (`scripts/ncu_traffic.sh`,")
print("raw CSV reduced by `scripts/make_tree_dram_json.py` into `profiles/tree_dram_bytes.json`; this table:")
print("`scripts/make_launch_list.py`).  ncu serialises the launches and runs them cold, so only the SHARES are compared with")
print("the CUDA-event times of the un-profiled bench run (`%s`, 4096 signals per step).\n" % sys.argv[1])
print("| kernel (events name) | launches | ncu ms at 1024 signals | ncu share | CUDA-event ms at 4096 signals | event share | DRAM MB / signal | DRAM TB/s (ncu) |")
print("|---|---|---|---|---|---|---|---|")
for k, (n, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    e = ev.get(k, 0.0)
    print("| %s | %d | %.3f | %.1f %% | %.2f | %.1f %% | %.2f | %.2f |" %
          (k, n, ms, 100 * ms / tot, e, 100 * e / tot_ev, by / 1e6, (by * 1024 / (ms * 1e-3)) / 1e12 if ms > 0.05 else 0))
print("\nTotal ncu %.2f ms per 1024 signals (x4 = %.1f ms) against %.1f ms of CUDA events per 4096 signals." % (tot, 4 * tot, tot_ev))
print("Physical DRAM traffic: tree %.2f MB / signal incl. the fused column kernel (algorithmic model of SURVEY 8(d): 61.87 MB)," %
      (d["tree_bytes_per_signal"] / 1e6))
print("remaining chirp-z kernels %.2f MB / signal." % (d["chirpz_bytes_per_signal"] / 1e6))

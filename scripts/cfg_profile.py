#!/usr/bin/env python
"""Per-kernel CUDA-event times of one batched call of config 3, 4, 5, 6 or 7 (see bench_configs.py):
python scripts/cfg_profile.py 3 [scale]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
import bench_configs as bc
cfg = int(sys.argv[1]); scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
if cfg in (3, 7):
    B = int(1024 * scale); Q, lam, G = bc.config3_inputs(B)
elif cfg == 4:
    B = int(2048 * scale); U = bc.config4_inputs(B)
else:
    B = int(1024 * scale); Q5 = bc.config5_inputs(B)
import fnft_b200 as F
L = F.lib(); L.fnft_errwarn_setprintf(None)
def run():
    if cfg == 3:
        o = F.nsev_default_opts(); o.bound_state_localization = F.BSLOC_NEWTON; o.discspec_type = F.DSTYPE_BOTH
        return F.nsev_batch(Q, (-20.0, 20.0), 0, None, 1, o, K=np.full(B, 8), Kmax=8, bound_states=G)
    if cfg == 7:
        o = F.nsev_default_opts(); o.discspec_type = F.DSTYPE_BOTH
        return F.nsev_batch(Q, (-20.0, 20.0), 4096, (-4.0, 4.0), 1, o, K=np.zeros(B), Kmax=64,
                            bound_states=np.zeros((B, 64), dtype=np.complex128))
    if cfg == 4:
        o = F.kdvv_default_opts(); o.discretization = F.KDV_4SPLIT4B
        return F.kdvv_batch(U, (-16.0, 15.0), 8192, (-3.55, 3.95), o)
    o = F.nsep_default_opts(); o.localization = (1 if cfg == 5 else 2); o.filtering = 1
    o.bounding_box[0], o.bounding_box[1], o.bounding_box[2], o.bounding_box[3] = -10, 10, -10, 10
    o.discretization = F.NSE_2SPLIT4B
    return F.nsep_batch(Q5, (0.0, 2 * np.pi), 4 * 4096, 4 * 4096, 1, o)
run()
L.fnft_b200_profile_enable(1)
run()
rep = L.fnft_b200_profile_report().decode()
L.fnft_b200_profile_enable(0)
tot = 0.0
for line in rep.strip().split("\n"):
    n, c, ms = line.split(); tot += float(ms)
    print("%-34s %4s launches %10.3f ms" % (n, c, float(ms)))
print("total kernel ms %.3f for B=%d" % (tot, B))

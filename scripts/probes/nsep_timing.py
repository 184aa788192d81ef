"""FNFT_B200_NSEP_TIMING=1 python scripts/probes/nsep_timing.py : host wall time of the phases of one config-5 call"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
import bench_configs as bc
import fnft_b200 as F
B = 1024
Q5 = bc.config5_inputs(B)
L = F.lib(); L.fnft_errwarn_setprintf(None)
def run():
    o = F.nsep_default_opts(); o.localization = 1; o.filtering = 1
    o.bounding_box[0], o.bounding_box[1], o.bounding_box[2], o.bounding_box[3] = -10, 10, -10, 10
    o.discretization = F.NSE_2SPLIT4B
    return F.nsep_batch(Q5, (0.0, 2 * np.pi), 4 * 4096, 4 * 4096, 1, o)
os.environ["X"] = "1"
run()
sys.stderr.write("==== second call\n"); sys.stderr.flush()
t0 = time.perf_counter(); run(); t1 = time.perf_counter()
sys.stderr.write("==== wall %.1f ms\n" % ((t1 - t0) * 1e3))

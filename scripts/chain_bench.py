#!/usr/bin/env python
"""fnft_kdvv with its default discretization (2SPLIT8B, degree 12) on a batch: wall time per call.
FNFT_B200_TREE_CONVERT=0 keeps the coefficient kernels for all levels.  python scripts/chain_bench.py [B] [D]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fnft_b200 as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
t = np.linspace(-16, 15, D)
u = np.stack([(0.5 + 0.01 * b) / np.cosh(t) ** 2 for b in range(B)])
F.kdvv_batch(u, [-16, 15], D, [-3.55, 3.95], None)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); ret, cs, rcs = F.kdvv_batch(u, [-16, 15], D, [-3.55, 3.95], None); ts.append(time.perf_counter() - t0)
print("kdvv default (2SPLIT8B) B=%d D=M=%d: %.1f ms per call, %.0f signals/s, ret %d, convert=%s" % (
    B, D, min(ts) * 1e3, B / min(ts), ret, os.environ.get("FNFT_B200_TREE_CONVERT", "1")))

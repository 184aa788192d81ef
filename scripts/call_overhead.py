#!/usr/bin/env python
"""Fixed cost of one fnft_nsev_batch call with pinned host buffers (config 2 shape) at small batch sizes."""
import ctypes as C, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, fnft_b200 as F
L = F.lib(); L.fnft_b200_set_device(0); torch.cuda.set_device(0)
T = np.array(bench.TT); XI = np.array(bench.XI)
opts = L.fnft_nsev_default_opts()
for B in (1, 16, 64, 128, 512, 4096):
    P = bench.signal_params(B)
    q = bench.signals_torch(P, 0, B, torch.device("cuda:0"))
    qh = torch.empty((B, bench.D), dtype=torch.complex128, pin_memory=True); qh.copy_(q)
    oh = torch.empty((B, bench.M), dtype=torch.complex128, pin_memory=True); oh.zero_()
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        rc = L.fnft_nsev_batch(B, bench.D, qh.data_ptr(), T.ctypes.data, bench.M, oh.data_ptr(), XI.ctypes.data, None, 0, None, None, 1, C.addressof(opts), None)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0
    print("B = %5d  median %.3f ms  min %.3f ms" % (B, sorted(ts[2:])[5], min(ts[2:])), flush=True)

#!/usr/bin/env python
"""Times fnft_nsev_batch with pinned host buffers for the BASELINE config-2 shape under the
current FNFT_B200_PIPE setting:  FNFT_B200_PIPE=8 python scripts/e2e_probe.py [B]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, fnft_b200 as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = F.lib()
L.fnft_b200_set_device(0)
torch.cuda.set_device(0)
P = bench.signal_params(B)
q = bench.signals_torch(P, B, torch.device("cuda:0"))
qh = torch.empty((B, bench.D), dtype=torch.complex128, pin_memory=True); qh.copy_(q)
oh = torch.empty((B, bench.M), dtype=torch.complex128, pin_memory=True)
T = np.array(bench.TT); XI = np.array(bench.XI)
opts = L.fnft_nsev_default_opts()
def step():
    rc = L.fnft_nsev_batch(B, bench.D, qh.data_ptr(), T.ctypes.data, bench.M, oh.data_ptr(), XI.ctypes.data,
                           None, 0, None, None, 1, C.addressof(opts), None)
    assert rc == 0, rc
for _ in range(2): step()
torch.cuda.synchronize()
ts = []
for _ in range(4):
    t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
print("PIPE=%s B=%d ms/step %s -> %.0f signals/s" % (os.environ.get("FNFT_B200_PIPE", "default"), B,
      ["%.1f" % t for t in ts], B / (min(ts) * 1e-3)))

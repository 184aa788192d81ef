cd /root/repo
FNFT_B200_PIPE_TRACE=1 python - <<'PY' 2> gpurun_out/trace.err
import ctypes as C, os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench, fnft_b200 as F
B = 4096
L = F.lib(); L.fnft_b200_set_device(0); torch.cuda.set_device(0)
P = bench.signal_params(B)
q = bench.signals_torch(P, 0, B, torch.device("cuda:0"))
qh = torch.empty((B, bench.D), dtype=torch.complex128, pin_memory=True); qh.copy_(q)
oh = torch.empty((B, bench.M), dtype=torch.complex128, pin_memory=True); oh.zero_()
T = np.array(bench.TT); XI = np.array(bench.XI)
opts = L.fnft_nsev_default_opts()
def step():
    rc = L.fnft_nsev_batch(B, bench.D, qh.data_ptr(), T.ctypes.data, bench.M, oh.data_ptr(), XI.ctypes.data, None, 0, None, None, 1, C.addressof(opts), None)
    assert rc == 0
ts=[]
for i in range(14):
    t0=time.perf_counter(); step(); ts.append((time.perf_counter()-t0)*1e3)
    sys.stderr.write("== step %d %.1f ms\n" % (i, ts[-1]))
print(["%.1f"%t for t in ts])
PY

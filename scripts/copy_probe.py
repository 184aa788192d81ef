#!/usr/bin/env python
"""Concurrent pinned H2D + D2H bandwidth of this box (1 GiB each way), for three kinds of host buffers:
torch pin_memory, cudaHostAlloc default, cudaHostAlloc write-combined (input only)."""
import ctypes as C, time, sys
import torch
rt = torch.cuda.cudart()
lib = C.CDLL("libcudart.so", mode=C.RTLD_GLOBAL) if False else None
n = 1 << 30
dev_in = torch.empty(n, dtype=torch.uint8, device="cuda")
dev_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cudart = C.CDLL(torch.__path__[0] + "/lib/libcudart.so.12") if False else None
import glob, os
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
cu = C.CDLL(cands[0])
def host_alloc(flags):
    p = C.c_void_p()
    rc = cu.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags))
    assert rc == 0, rc
    C.memset(p, 1, n)
    return p
def run(name, pin, pout):
    cu.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    best = {}
    for mode in ("h2d", "d2h", "both"):
        ts = []
        for it in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode in ("h2d", "both"):
                cu.cudaMemcpyAsync(C.c_void_p(dev_in.data_ptr()), pin, n, 1, C.c_void_p(s1.cuda_stream))
            if mode in ("d2h", "both"):
                cu.cudaMemcpyAsync(pout, C.c_void_p(dev_out.data_ptr()), n, 2, C.c_void_p(s2.cuda_stream))
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        t = min(ts[1:])
        best[mode] = (n * (2 if mode == "both" else 1)) / t / 1e9
    print("%-28s h2d %.1f GB/s  d2h %.1f GB/s  both %.1f GB/s (sum)" % (name, best["h2d"], best["d2h"], best["both"]), flush=True)
tp_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); tp_in.fill_(1)
tp_out = torch.empty(n, dtype=torch.uint8, pin_memory=True); tp_out.fill_(0)
run("torch pin_memory", C.c_void_p(tp_in.data_ptr()), C.c_void_p(tp_out.data_ptr()))
a = host_alloc(0); b = host_alloc(0)
run("cudaHostAlloc default", a, b)
w = host_alloc(4)  # cudaHostAllocWriteCombined
run("write-combined input", w, b)

"""Ad-hoc GPU sanity run: product library vs oracle/_ref on a few cases + timing.
Usage (on a GPU box):  python scripts/gpu_check.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fnft_b200 as F
from oracle import ref_lib as R

def relerr(a, b):
    return np.abs(a - b).sum() / np.abs(b).sum()

print("devices", F.device_count())
rng = np.random.default_rng(0)
# 1. fscatter
for disc in [F.NSE_2SPLIT4B, F.NSE_2SPLIT2A]:
    for D in [2, 3, 8, 100, 256, 1000, 1024, 4096, 5000, 16384]:
        t = np.linspace(-10, 10, D); eps = 20.0 / (D - 1)
        q = 2.2 / np.cosh(t) * np.exp(0.5j * t)
        r0, tm0, d0, W0 = R.nse_fscatter(q, eps, 1, disc)
        r1, tm1, d1, W1 = F.nse_fscatter(q, eps, 1, disc)
        e = max(relerr(tm1[k] * 2.0 ** W1, tm0[k] * 2.0 ** W0) for k in range(4))
        print("fscatter disc", disc, "D", D, "rc", r0, r1, "deg", d0, d1, "W", W0, W1, "err %.2e" % e)
# 2. chirpz
p = rng.standard_normal(4) + 1j * rng.standard_normal(4)
for (A, W, M) in [(0.95, np.exp(0.3j), 6)]:
    print("chirpz", relerr(F.poly_chirpz(p, A, W, M)[1], R.poly_chirpz(p, A, W, M)[1]))
# 3. nsev contspec
for D, M in [(256, 8), (1024, 1024), (4096, 4096), (16384, 16384)]:
    t = np.linspace(-32, 32, D)
    q = 5.4 / np.cosh(t) * np.exp(-6j * t)
    for cst in [0, 2]:
        o0 = R.nsev_default_opts(); o0.contspec_type = cst
        o1 = F.nsev_default_opts(); o1.contspec_type = cst
        t0 = time.time(); r0, c0, *_ = R.nsev(q, [-32, 32], M, [-10, 10], 1, o0); t_ref = time.time() - t0
        t0 = time.time(); r1, c1, *_ = F.nsev(q, [-32, 32], M, [-10, 10], 1, o1); t_gpu = time.time() - t0
        errs = [relerr(c1[i * M:(i + 1) * M], c0[i * M:(i + 1) * M]) for i in range(len(c0) // M)]
        print("nsev D", D, "cstype", cst, "rc", r0, r1, "err", ["%.2e" % e for e in errs], "t_ref %.3f t_gpu %.3f" % (t_ref, t_gpu))
# 4. kdvv
for D, M in [(256, 64), (8192, 8192)]:
    t = np.linspace(-16, 15, D); u = 2.0 / np.cosh(t) ** 2
    for disc in [F.KDV_2SPLIT4B, F.KDV_4SPLIT4B, F.KDV_2SPLIT2A]:
        o0 = R.lib().fnft_kdvv_default_opts(); o0.discretization = disc
        o1 = F.kdvv_default_opts(); o1.discretization = disc
        r0, c0 = R.kdvv(u, [-16, 15], M, [-3.55, 3.95], o0)
        r1, c1 = F.kdvv(u, [-16, 15], M, [-3.55, 3.95], o1)
        print("kdvv D", D, "disc", disc, r0, r1, "err %.2e" % relerr(c1, c0))
# 5. bound states (sech with 3 bound states: 0.5i,1.5i,2.5i for A=3)
D = 2048; t = np.linspace(-16, 16, D); q = 3.0 / np.cosh(t)
for disc, nm in [(F.NSE_2SPLIT4B, "2SPLIT4B"), (F.NSE_4SPLIT4B, "4SPLIT4B")]:
    o0 = R.nsev_default_opts(); o0.bound_state_localization = 1; o0.discspec_type = 2; o0.discretization = disc
    o1 = F.nsev_default_opts(); o1.bound_state_localization = 1; o1.discspec_type = 2; o1.discretization = disc
    g = np.array([0.45j + 0.01, 1.52j - 0.02, 2.48j, 0.46j])
    r0, c0, K0, b0, n0 = R.nsev(q, [-16, 16], 16, [-2, 2], 1, o0, K=4, bound_states=g)
    r1, c1, K1, b1, n1 = F.nsev(q, [-16, 16], 16, [-2, 2], 1, o1, K=4, bound_states=g)
    print(nm, "bound states rc", r0, r1, "K", K0, K1)
    print("  ref", b0, n0[:2 * K0]); print("  gpu", b1, n1[:2 * K1])
    print("  contspec err %.2e" % relerr(c1, c0))
# 6. throughput, config 2 shape, small batch
D = M = 16384; B = 64
t = np.linspace(-32, 32, D)
A = rng.uniform(0.5, 5.4, B)[:, None]; l0 = rng.uniform(-3, 3, B)[:, None]
Q = A / np.cosh(t)[None, :] * np.exp(-2j * l0 * t[None, :])
for it in range(3):
    t0 = time.time(); ret, cs, *_ = F.nsev_batch(Q, [-32, 32], M, [-10, 10], 1, None); dt = time.time() - t0
    print("batch", B, "ret", ret, "time %.3f s -> %.1f signals/s (host buffers, pageable)" % (dt, B / dt), "launches", F.launch_count())
r0, c0, *_ = R.nsev(Q[3], [-32, 32], M, [-10, 10], 1, None)
print("batch parity signal 3: %.2e" % relerr(cs[3], c0))

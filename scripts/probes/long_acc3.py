import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fnft_b200 as F
from oracle import fnft_oracle as O, ref_lib as R
F.lib().fnft_errwarn_setprintf(None)
def truth_ab(q, T, XI, M, idx):
    D = len(q); eps_t = (T[1] - T[0]) / (D - 1)
    xi = (XI[0] + (XI[1] - XI[0]) / (M - 1) * idx).astype(np.longdouble)
    z = np.exp(1j * xi * np.longdouble(eps_t)); z2 = z * z
    P = O.akns_leaves(q, -np.conj(q), eps_t, O.AKNS_2SPLIT4B).astype(np.clongdouble)
    v1, v2 = np.ones(len(idx), dtype=np.clongdouble), np.zeros(len(idx), dtype=np.clongdouble)
    for k in range(D - 1, -1, -1):
        m11 = P[0, k, 0] * z2 + P[0, k, 1] * z + P[0, k, 2]; m12 = P[1, k, 0] * z2 + P[1, k, 1] * z + P[1, k, 2]
        m21 = P[2, k, 0] * z2 + P[2, k, 1] * z + P[2, k, 2]; m22 = P[3, k, 0] * z2 + P[3, k, 1] * z + P[3, k, 2]
        v1, v2 = m11 * v1 + m12 * v2, m21 * v1 + m22 * v2
    L = np.longdouble
    ph_a = -L(eps_t) * D + (L(T[1]) + L(eps_t) / 2) - (L(T[0]) - L(eps_t) / 2)
    ph_b = -L(eps_t) * D - (L(T[1]) + L(eps_t) / 2) - (L(T[0]) - L(eps_t) / 2)
    return (v1 * np.exp(1j * xi * ph_a)), (v2 * np.exp(1j * xi * ph_b))
XI, M = (-6.0, 6.0), 96
idx = np.arange(0, M, 8)
for T, D in (((0.0005, 40.0), 75001), ((-40.0, 40.0), 100000), ((-40.0, 40.0), 16385)):
    rng = np.random.default_rng(33); t = np.linspace(T[0], T[1], D)
    q = 1.7 / np.cosh(t / 1.3 - 0.4) * np.exp(0.9j * t + 1j) + 0.02 * (rng.standard_normal(D) + 1j * rng.standard_normal(D))
    o = F.nsev_default_opts(); o.contspec_type = 2
    ta, tb = truth_ab(q, T, XI, M, idx)
    for name, lib in (("ours", F), ("ref", R if R.available() else None)):
        if lib is None: continue
        if name == "ref":
            o = R.nsev_default_opts(); o.contspec_type = 2
            ret, cs, *_ = R.nsev(q, np.array(T), M, np.array(XI), 1, o)
        else:
            ret, cs, *_ = F.nsev(q, T, M, XI, 1, o)
        a, b = cs[M:2*M][idx], cs[2*M:][idx]
        ea = (a.astype(np.clongdouble) / ta - 1).astype(np.complex128); eb = (b.astype(np.clongdouble) / tb - 1).astype(np.complex128)
        print(name, T, D)
        for j in range(len(idx)):
            print("   m %2d |a| %.3f |b| %.2e  a/a_true-1 = %+.2e %+.2ei   b/b_true-1 = %+.2e %+.2ei" % (idx[j], abs(a[j]), abs(b[j]), ea[j].real, ea[j].imag, eb[j].real, eb[j].imag))

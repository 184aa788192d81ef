#!/usr/bin/env python
"""Accuracy of the continuous spectrum against a long-double product of the per-sample leaf matrices
(no FFT products, no chirp-z) as the signal grows: ours (one tree up to D = 131072, segmented beyond), and the
reference when oracle/_ref is present.  Prints the two figures of tests/common.py::parity_contract (x 1e-9)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fnft_b200 as F
from common import parity_contract
from oracle import fnft_oracle as O, ref_lib as R

def truth_rho(q, T, XI, M, idx):
    D = len(q); eps_t = (T[1] - T[0]) / (D - 1)
    xi = (XI[0] + (XI[1] - XI[0]) / (M - 1) * idx).astype(np.longdouble)
    z = np.exp(1j * xi * np.longdouble(eps_t)); z2 = z * z
    P = O.akns_leaves(q, -np.conj(q), eps_t, O.AKNS_2SPLIT4B).astype(np.clongdouble)
    v1, v2 = np.ones(len(idx), dtype=np.clongdouble), np.zeros(len(idx), dtype=np.clongdouble)
    for k in range(D - 1, -1, -1):
        m11 = P[0, k, 0] * z2 + P[0, k, 1] * z + P[0, k, 2]; m12 = P[1, k, 0] * z2 + P[1, k, 1] * z + P[1, k, 2]
        m21 = P[2, k, 0] * z2 + P[2, k, 1] * z + P[2, k, 2]; m22 = P[3, k, 0] * z2 + P[3, k, 1] * z + P[3, k, 2]
        v1, v2 = m11 * v1 + m12 * v2, m21 * v1 + m22 * v2
    ph = np.longdouble(-2.0 * (T[1] + 0.5 * eps_t))
    return (v2 / v1 * np.exp(1j * xi * ph)).astype(np.complex128)

F.lib().fnft_errwarn_setprintf(None)
T, XI, M = (-40.0, 40.0), (-6.0, 6.0), 96
idx = np.arange(0, M, 4)
for noise in (0.0, 0.02):
    for D in (131072, 150001, 300000, 1000003):
        rng = np.random.default_rng(33); t = np.linspace(T[0], T[1], D)
        q = 1.7 / np.cosh(t / 1.3 - 0.4) * np.exp(0.9j * t + 1j) + noise * (rng.standard_normal(D) + 1j * rng.standard_normal(D))
        ret, cs, *_ = F.nsev(q, T, M, XI, 1, None)
        assert ret == 0, ret
        tr = truth_rho(q, T, XI, M, idx)
        line = "noise %.2f D %6d max|rho| %.2e ours vs truth %.3f %.3f" % ((noise, D, np.abs(tr).max()) + parity_contract(cs[idx], tr))
        if R.available() and D <= 300000:
            rr, ref, *_ = R.nsev(q, np.array(T), M, np.array(XI), 1)
            line += "   reference vs truth %.3f %.3f" % parity_contract(ref[idx], tr)
        print(line, flush=True)

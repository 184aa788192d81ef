import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import fnft_b200 as F
from common import parity_contract
sys.argv = sys.argv[:1]
import importlib.util
spec = importlib.util.spec_from_file_location("la", os.path.join(ROOT, "scripts", "long_accuracy.py"))
src = open(os.path.join(ROOT, "scripts", "long_accuracy.py")).read().split("F.lib().fnft_errwarn_setprintf")[0]
exec(src)
F.lib().fnft_errwarn_setprintf(None)
XI, M = (-6.0, 6.0), 96
idx = np.arange(0, M, 4)
for T in ((-40.0, 40.0), (-40.0, 0.0), (0.0005, 40.0)):
    for D in (65000, 70000, 75000, 75001, 80000, 90000, 100000):
        rng = np.random.default_rng(33); t = np.linspace(T[0], T[1], D)
        q = 1.7 / np.cosh(t / 1.3 - 0.4) * np.exp(0.9j * t + 1j) + 0.02 * (rng.standard_normal(D) + 1j * rng.standard_normal(D))
        o = F.nsev_default_opts(); o.contspec_type = 2
        ret, cs, *_ = F.nsev(q, T, M, XI, 1, o)
        tr = truth_rho(q, T, XI, M, idx)
        a, b = cs[M:2*M], cs[2*M:]
        print("T", T, "D", D, "rho vs truth %.3f %.3f" % parity_contract(cs[:M][idx], tr), " | |a|^2+|b|^2-1| max %.2e" % np.abs(np.abs(a)**2+np.abs(b)**2-1).max(), flush=True)

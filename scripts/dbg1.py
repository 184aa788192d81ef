import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fnft_b200 as F
from oracle import fnft_oracle as O, ref_lib as R
D,M,kappa=126,40,-1
T, XI = (-32.0, 32.0), (-10.0, 10.0)
t = np.linspace(T[0], T[1], D)
q = 1.7 / np.cosh(t) * np.exp(-3j * t + 0.4j * np.sin(t))
o = F.nsev_default_opts(); o.contspec_type = 2
ret, cs, *_ = F.nsev(q, T, M, XI, kappa, o)
ref = O.nsev_contspec(q, T, M, XI, kappa, 11, cstype=2)
o0=R.nsev_default_opts(); o0.contspec_type=2
r0, c0, *_ = R.nsev(q, T, M, XI, kappa, o0)
for part in range(3):
    a=cs[part*M:(part+1)*M]; b=ref[part*M:(part+1)*M]; c=c0[part*M:(part+1)*M]
    print(part, "ours-oracle max rel", np.max(np.abs(a-b)/np.abs(b)), "ours-ref", np.max(np.abs(a-c)/np.abs(c)), "oracle-ref", np.max(np.abs(b-c)/np.abs(c)), "max|.|", np.abs(c).max(), "min", np.abs(c).min())

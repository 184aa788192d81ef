"""Golden vectors for fnft_nsev with the discretizations ES4 and TES4 (fnft_nse_discretization_t 26, 27: one / three
Pauli-expanded matrix exponentials per step on (q, q', q''), finite-difference preprocessing), produced by the UNMODIFIED
reference compiled into oracle/_ref (oracle/ref_lib.py) -- run in the build container:
    python tests/golden/make_golden_es4.py
Same layout as golden_cf4_3.npz.  D = 255 / 100 / 300 exercise lengths that are not powers of two; the Richardson runs
(src/fnft_nsev.c:316-442) exercise the sub-sampled preprocessing (nskip_per_step = 2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_lib as R  # noqa: E402

G = {}
for disc, D, kappa in [(d, D, k) for d in (26, 27) for D, k in ((100, 1), (256, 1), (255, 1), (256, -1), (300, -1))]:
    tt = np.linspace(-10, 10, D)
    qs = 2.7 / np.cosh(tt) * np.exp(0.4j * tt)
    o = R.nsev_default_opts()
    o.discretization, o.bound_state_localization, o.discspec_type, o.contspec_type = disc, 1, 2, 2
    gs = np.array([0.2j - 0.2, 1.2j - 0.21, 2.19j - 0.2])
    ret, cs, K, bs, nc = R.nsev(qs, [-10, 10], 20, [-2, 2.5], kappa, o, K=3, bound_states=gs)
    assert ret == 0, (disc, D, kappa, ret)
    key = f"refrun/slow/{disc}/{D}/{kappa}"
    G[key + "/q"] = qs
    G[key + "/guesses"] = gs
    G[key + "/cs"] = cs
    G[key + "/bs"] = bs[:K]
    G[key + "/nc"] = nc[:2 * K]
    if D == 256:
        o.richardson_extrapolation_flag = 1
        ret, cs, K, bs, nc = R.nsev(qs, [-10, 10], 20, [-2, 2.5], kappa, o, K=3, bound_states=gs)
        assert ret == 0
        key = f"refrun/slow_richardson/{disc}/{D}/{kappa}"
        G[key + "/cs"] = cs
        G[key + "/bs"] = bs[:K]
        G[key + "/nc"] = nc[:2 * K]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_es4.npz"), **G)
print("wrote", len(G), "arrays")

#!/usr/bin/env python
"""Per-scheme parity report (GPU box): ours vs recorded reference vs numpy oracle for every
polynomial discretization (tests/golden refrun/schemes_*).  Prints the two parity-contract figures
(L1-relative, pointwise with floor) in units of 1e-9."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fnft_b200 as F
from common import parity_contract
from oracle import fnft_oracle as O
g = np.load(os.path.join(ROOT, "tests/golden/golden.npz"))
F.lib().fnft_errwarn_setprintf(None)
fmt = lambda t: "(%.3g, %.3g)" % tuple(t)
for k in sorted(g.files):
    if k.startswith("refrun/schemes_nsev/") and k.endswith("/q"):
        _, _, disc, kappa, _ = k.split("/")
        disc, kappa = int(disc), int(kappa)
        o = F.nsev_default_opts(); o.discretization = disc; o.contspec_type = F.CSTYPE_BOTH
        ret, cs, *_ = F.nsev(g[k], [-6, 6], 24, [-2.5, 3.25], kappa, o)
        ref = g[k[:-1] + "cs"]
        orc = O.nsev_contspec(g[k], [-6, 6], 24, [-2.5, 3.25], kappa, disc, cstype=2)
        for part, nm in enumerate(("rho", "a", "b")):
            sl = slice(part * 24, (part + 1) * 24)
            print("nsev disc %2d kappa %+d %-3s ours-ref %s ours-oracle %s oracle-ref %s" % (
                disc, kappa, nm, fmt(parity_contract(cs[sl], ref[sl])), fmt(parity_contract(cs[sl], orc[sl])),
                fmt(parity_contract(orc[sl], ref[sl]))))

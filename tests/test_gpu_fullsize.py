"""Parity at the FULL sizes of BASELINE.json configs 3, 4, 5 and 7: the CUDA path
through the C-ABI against the unmodified reference (oracle/_ref, built by oracle/Makefile, travels to the GPU box)
on the same seeded inputs that bench.py uses.  The reference runs on the host cores of the box in a process pool,
so the sample sizes are chosen for a few seconds each.

Config 5 additionally settles WHOSE error the 1.7e-7 difference at D = 4096 is: both implementations' roots are
put into a long-double evaluation of the polynomial the grid search works on (oracle restatement of
src/fnft_nsep.c:222-436); the reference's chirp-z samples carry the absolute cpow floor, ours do not.
"""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from common import ensure_lib, parity_contract, parity_pointwise  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    ensure_lib()
    import fnft_b200 as F_
    F_.quiet(True)
    return F_


@pytest.fixture(scope="module")
def R():
    from oracle import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref/libfnft_ref.so not built")
    return ref_lib


@pytest.fixture(scope="module")
def BM():
    import bench
    return bench


def _pool(fn, tasks):
    ctx = mp.get_context("spawn")  # CUDA may already be initialised in this process
    with ctx.Pool(min(len(tasks), max(1, (os.cpu_count() or 2) - 1))) as p:
        return p.map(fn, tasks)


@pytest.fixture(scope="module")
def solitons(BM, R):
    """64 of config 3's 8-soliton signals (D = 4096), synthesised by the reference's fnft_nsev_inverse"""
    n = 64
    lams, bn, guesses = BM.config3_params()
    idx = BM.spread(BM.C3["B"], n)
    q = np.stack(_pool(BM._soliton_worker, [(lams[i], bn[i], BM.C3["D"], BM.C3["T"]) for i in idx]))
    return idx, q, guesses[idx]


def test_config3_newton_bound_states_full_size(F, BM, solitons):
    # D = 4096, K = 8 guesses, Newton (niter 10), bsfilt FULL, dstype BOTH, 64 signals in one batched call
    idx, q, g = solitons
    K = BM.C3["K"]
    ref = _pool(BM._ref3, [(q[i], g[i]) for i in range(len(idx))])
    o = F.nsev_default_opts()
    o.bound_state_localization = 1
    o.discspec_type = 2
    ret, _, Ka, bs, nc, rcs = F.nsev_batch(q, BM.C3["T"], 0, None, 1, o, K=np.full(len(idx), K), Kmax=K, bound_states=g)
    assert ret == 0 and (rcs == 0).all()
    worst = 0.0
    for i in range(len(idx)):
        assert ref[i][0] == 0
        assert int(Ka[i]) == int(ref[i][1]) == K, (i, Ka[i], ref[i][1])
        worst = max(worst, BM.compare3(Ka[i], bs[i], nc[i], ref[i], K))
    # |d lambda| <= 1e-9 |lambda|, |d b| <= 1e-9 |b|, |d res| <= 1e-9 |res| per eigenvalue
    assert worst <= 1e-9, worst


def test_config4_kdvv_4split4b_full_size(F, BM, R):
    # fnft_kdvv, 4SPLIT4B, D = M = 8192: 16 signals spread over the batch of 2048
    n = 16
    idx = BM.spread(BM.C4["B"], n)
    u = np.stack([BM.config4_inputs(i, i + 1)[0] for i in idx])
    ref = _pool(BM._ref4, [(u[i],) for i in range(n)])
    o = F.kdvv_default_opts()
    o.discretization = 19
    ret, cs, rcs = F.kdvv_batch(u, BM.C4["T"], BM.C4["M"], BM.C4["XI"], o)
    assert ret == 0 and (rcs == 0).all()
    for i in range(n):
        assert ref[i][0] == 0
        e1, e2 = parity_contract(cs[i], ref[i][1])
        assert e1 < 1 and e2 < 1, (i, e1, e2)
        # SURVEY 8(c)(ii)/(iii) as written: pointwise 1e-9 relative where |ref| >= 1e-6 max, absolute below
        assert parity_pointwise(cs[i], ref[i][1]) < 1, i


def _horner_ld(p, z):
    r = np.zeros(z.shape, dtype=np.clongdouble)
    for c in p.astype(np.clongdouble):
        r = r * z + c
    return r


def _ideal_gridsearch(O, p, PHI, M, lam, eps_t, deg0):
    """The grid-search root estimates (src/private/fnft__poly_roots_fftgridsearch.c:78-148) recomputed from
    LONG-DOUBLE samples of p at the windows that produced the points lam: what the algorithm returns when the
    three chirp-z transforms are exact."""
    ld = np.longdouble
    eps = (ld(PHI[1]) - ld(PHI[0])) / (M - 1)
    ang = np.real(lam).astype(ld) * ld(eps_t) * 2 / deg0          # z = exp(2i lambda eps_t / deg0)
    i0 = np.rint((ang - ld(PHI[0])) / eps).astype(np.int64)
    best = np.full(lam.shape, np.nan + 0j, dtype=np.clongdouble)
    for di in (0, -1, 1):                                           # the window is the one next to the root
        i = np.clip(i0 + di, 1, M - 2)
        zr = O.fftgridsearch_windows(p, PHI, M, i, _horner_ld)
        lr = np.log(zr) * deg0 / (2j * ld(eps_t))                   # z_to_lambda, fnft__akns_discretization.c:225-240
        take = ~np.isnan(zr) & (np.isnan(best) | (np.abs(lr - lam) < np.abs(best - lam)))
        best = np.where(take, lr, best)
    return best


def test_config5_nsep_gridsearch_full_size_and_whose_error_it_is(F, BM, R):
    # fnft_nsep, 2SPLIT4B, D = 4096, grid search, manual box [-10, 10]^2.  Same numbers of points as the
    # reference; positions differ by up to ~2e-7.  Both implementations run the SAME estimator on samples of the
    # same polynomials; evaluated on long-double samples that estimator has one well-defined output per window.
    # Ours must sit on it, the reference (chirp-z with cpow(W, n^2/2), absolute error floor ~1e-11 max|p|,
    # DESIGN.md 5) is the one that is off.
    from oracle import fnft_oracle as O
    idx = np.array([0, 10, 614])             # 10: bench.py's worst case (1.7e-7 relative); 614: survey below
    n = len(idx)
    q = np.stack([BM.config5_inputs(i, i + 1)[0] for i in idx])
    ref = _pool(BM._ref5, [(q[i],) for i in range(n)])
    o = F.nsep_default_opts()
    o.localization = 1
    o.filtering = 1
    o.bounding_box[0], o.bounding_box[1], o.bounding_box[2], o.bounding_box[3] = -10, 10, -10, 10
    o.discretization = 11
    Kmax = max(len(r[1]) for r in ref) + 16
    Mmax = max(len(r[2]) for r in ref) + 16
    ret, Ka, main, Ma, aux, rcs = F.nsep_batch(q, BM.C5["T"], Kmax, Mmax, 1, o)
    assert ret == 0 and (rcs == 0).all()
    deg0 = 2
    worst = dict(ours_aux=0.0, ref_aux=0.0, ours_main=0.0, ref_main=0.0, diff=0.0)
    for i in range(n):
        assert ref[i][0] == 0
        m0, a0 = ref[i][1], ref[i][2]
        m1, a1 = main[i][:int(Ka[i])], aux[i][:int(Ma[i])]
        assert len(m0) == len(m1) and len(a0) == len(a1), (i, len(m0), len(m1), len(a0), len(a1))
        di = BM.compare5(Ka[i], main[i], Ma[i], aux[i], ref[i])
        worst["diff"] = max(worst["diff"], di)
        pp, pm, paux, eps_t, deg = O.nsep_gridsearch_polys(q[i], BM.C5["T"], +1, 11)
        PHI = (eps_t * -10.0 * 2 / deg0, eps_t * 10.0 * 2 / deg0)   # src/fnft_nsep.c:287-289
        M = 32 * deg                                                # oversampling_factor * deg, :265, :324
        ideal = _ideal_gridsearch(O, paux, PHI, M, a1, eps_t, deg0)
        assert not np.isnan(ideal).any()
        worst["ours_aux"] = max(worst["ours_aux"], float(np.abs(a1 - ideal).max()))
        worst["ref_aux"] = max(worst["ref_aux"], float(np.abs(a0 - ideal).max()))
        # main spectrum: the points of p+ come first, then those of p- (:318-392); a point belongs to the
        # polynomial whose ideal estimate is next to it
        ip = _ideal_gridsearch(O, pp, PHI, M, m1, eps_t, deg0)
        im = _ideal_gridsearch(O, pm, PHI, M, m1, eps_t, deg0)
        dp, dm = np.abs(m1 - ip), np.abs(m1 - im)
        idl = np.where(np.isnan(dm) | (dp <= dm), ip, im)
        assert not np.isnan(idl).any()
        worst["ours_main"] = max(worst["ours_main"], float(np.abs(m1 - idl).max()))
        worst["ref_main"] = max(worst["ref_main"], float(np.abs(m0 - idl).max()))
        print("signal", int(idx[i]), "ours-ref %.2e" % di, "| aux: ours-ideal %.2e ref-ideal %.2e" %
              (float(np.abs(a1 - ideal).max()), float(np.abs(a0 - ideal).max())),
              "| main: ours-ideal %.2e ref-ideal %.2e" % (float(np.abs(m1 - idl).max()), float(np.abs(m0 - idl).max())))
    print("config 5, |lambda - ideal estimator output|:", worst)
    # Signal 10 is the one behind the 1.7e-7 (relative to |lambda| <= 10) that bench.py's parity gate reports for
    # config 5: the REFERENCE is 1.66e-6 away from the idealised estimator there (CPU-only survey of the first 16
    # signals: all others <= 3.7e-8).  Survey of 16 more signals with the GPU (session log, one B200 box): |ours - ref| <= 4.5e-9 everywhere; auxiliary spectrum:
    # both 5.9e-9 from the idealised estimator and 1e-13 from each other (the idealisation puts the samples exactly
    # on the rings, the implementations put them where the rounded W of the reference puts them; d lambda =
    # d z / eps_t amplifies 1e-11 in z 650 times); main spectrum: ours 3.0e-9 on 12 signals where the reference is
    # 3e-9 ... 4.5e-8 (signal 614: 2.99e-9 vs 4.54e-8, signal 136: 2.98e-9 vs 1.15e-8), and 1.4 - 2.5e-8 for both on
    # the four signals with near-double points (273, 409, 887, 1023), where the two agree to 2e-9.
    assert worst["ours_aux"] <= 1e-8 and abs(worst["ours_aux"] - worst["ref_aux"]) <= 1e-10, worst
    assert worst["ours_main"] <= 3e-8, worst             # signal 10: 1.86e-8 (near-double points), others 3.0e-9
    assert worst["ref_main"] >= 50 * worst["ours_main"], worst    # ... the reference is the one that is off: 1.66e-6
    assert worst["diff"] <= worst["ref_main"] + worst["ours_main"], worst


def _ref7(args):
    """the reference with its DEFAULT options (SUBSAMPLE_AND_REFINE, eiscor replaced by the LAPACK companion-matrix
    shim of oracle/eiscor_shim.c) plus residues, on one signal"""
    # one BLAS thread per worker: the shim's LAPACK call otherwise starts a thread per core in every one of the
    # worker processes and the box thrashes (9 minutes instead of 10 seconds)
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "SCIPY_OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"
    sys.path.insert(0, ROOT)
    from oracle import ref_lib as Rl
    q, T, kmax = args
    Rl.lib().fnft_errwarn_setprintf(None)
    o = Rl.nsev_default_opts()
    o.discspec_type = 2
    ret, cs, K, bs, nc = Rl.nsev(q, T, 0, None, 1, o, K=kmax, want_contspec=False)
    return ret, K, bs[:K].copy(), nc[:2 * K].copy()


def test_config7_default_options_find_what_the_reference_finds(F, BM, solitons):
    # fnft_nsev with its default options (bsloc SUBSAMPLE_AND_REFINE: roots of the sub-sampled a(z) by the GPU
    # Aberth-Ehrlich finder instead of eiscor, Newton refinement on the full signal) on 12 of config 3's 8-soliton
    # signals, D = 4096: the same NUMBER of bound states as the reference on every signal and the same eigenvalues,
    # norming constants and residues to 1e-9 (Newton on the same recurrence converges to the same zeros whatever
    # root finder supplied the start values).
    idx, q, _ = solitons
    n, kmax = 12, 64
    q = q[:n]
    ref = _pool(_ref7, [(q[i], BM.C3["T"], kmax) for i in range(n)])
    o = F.nsev_default_opts()
    o.discspec_type = 2
    ret, _, Ka, bs, nc, rcs = F.nsev_batch(q, BM.C3["T"], 0, None, 1, o, K=np.full(n, kmax), Kmax=kmax,
                                           bound_states=np.zeros((n, kmax), dtype=np.complex128))
    assert ret == 0 and (rcs == 0).all()
    missing, worst = [], 0.0
    for i in range(n):
        assert ref[i][0] == 0
        Kr, bsr, ncr = ref[i][1], ref[i][2], ref[i][3]
        if int(Ka[i]) != Kr:
            missing.append((int(idx[i]), int(Ka[i]), Kr))
            continue
        for j in range(Kr):
            jj = int(np.argmin(np.abs(bs[i][:Kr] - bsr[j])))
            worst = max(worst, abs(bs[i][jj] - bsr[j]) / abs(bsr[j]),
                        abs(nc[i][jj] - ncr[j]) / abs(ncr[j]),
                        abs(nc[i][Kr + jj] - ncr[Kr + j]) / abs(ncr[Kr + j]))
    print("config 7: signals with a different number of bound states:", missing, "worst relative error:", worst)
    assert not missing, missing
    assert worst <= 1e-9, worst

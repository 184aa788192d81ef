"""GPU tests added in round 2 (run with -m gpu): the advisor's findings of round 1, devices other
than 0, the several-GPUs-behind-one-call fan-out, and BASELINE-size parity against the live
reference (oracle/_ref) for the configurations bench.py reports."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from common import parity_contract, rel_err, sech_chirp
from oracle import fnft_oracle as O
from oracle import ref_lib as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def F():
    import fnft_b200
    if fnft_b200.device_count() < 1:
        pytest.fail("no CUDA device visible to libfnft_b200.so (there is no CPU fallback to test)")
    fnft_b200.lib().fnft_errwarn_setprintf(None)
    return fnft_b200


def _signals(B, D, T, seed):
    rng = np.random.default_rng(seed)
    t = np.linspace(T[0], T[1], D)
    return np.stack([rng.uniform(0.6, 3.0) / np.cosh(t / rng.uniform(0.6, 1.6)) *
                     np.exp(1j * rng.uniform(-2, 2) * t + 1j * rng.uniform(0, 6.28)) for _ in range(B)])


# ------------------------------------------------------------------ advisor, round 1
@pytest.mark.parametrize("D", [2048, 4096])
def test_fast_eigenvalue_together_with_contspec_long_signal(F, D):
    """bsloc_FAST_EIGENVALUE reuses the transfer matrix of the continuous-spectrum pass.  From
    deg*D + M > 4096 on that pass reads (a, b) straight from the level buffer and used to skip the
    kernel that writes the matrix (round-1 advisor finding): the roots then came from stale memory."""
    T, XI, M = (-12.0, 12.0), (-3.0, 3.0), 300
    q = 2.6 / np.cosh(np.linspace(T[0], T[1], D)) * np.exp(0.4j * np.linspace(T[0], T[1], D))
    o = F.nsev_default_opts()
    o.bound_state_localization = F.BSLOC_FAST_EIGENVALUE
    Kmax = 64
    r_only, _, K0, bs0, nc0 = F.nsev(q, T, 0, None, 1, o, K=Kmax, want_contspec=False)
    # a call on other data in between, so that stale workspace contents cannot look right
    F.nsev(_signals(1, D, T, 5)[0], T, M, XI, 1, o, K=Kmax)
    r_both, cs, K1, bs1, nc1 = F.nsev(q, T, M, XI, 1, o, K=Kmax)
    assert r_only == 0 and r_both == 0
    assert K0 == K1 and K0 >= 2, (K0, K1)
    assert np.abs(np.sort_complex(bs0) - np.sort_complex(bs1)).max() < 1e-12
    # roots of a(z): one Newton step in long double on the oracle's polynomial does not move them
    eps_t = (T[1] - T[0]) / (D - 1)
    tm, _, _ = O.nse_fscatter(q, eps_t, 1, O.NSE_2SPLIT4B)
    z = np.exp(2j * bs1.astype(np.clongdouble) * eps_t / 2)
    pv, dv = np.zeros(K1, dtype=np.clongdouble), np.zeros(K1, dtype=np.clongdouble)
    for ck in tm[0].astype(np.clongdouble):
        dv = dv * z + pv
        pv = pv * z + ck
    assert (np.abs(pv / dv) <= 1e-11 * np.abs(z)).all(), np.abs(pv / dv)
    ref = O.nsev_contspec(q, T, M, XI, 1)
    assert max(parity_contract(cs, ref)) < 1


def test_reference_convention_kmax_does_not_blow_up_the_workspace(F):
    """Kmax = fnft_nsev_max_K(D) per signal (the reference's convention) on a batch: the scratch of
    the norming-constant kernel is sized by the eigenvalues present, not by B*Kmax."""
    B, D, T = 96, 1024, (-10.0, 10.0)
    Q = _signals(B, D, T, 11)
    o = F.nsev_default_opts()
    o.bound_state_localization = F.BSLOC_NEWTON
    o.discspec_type = F.DSTYPE_BOTH
    Kmax = int(F.lib().fnft_nsev_max_K(D, C.addressof(o)))
    assert Kmax == 2 * D
    G = np.zeros((B, 3), dtype=np.complex128)
    G[:] = [0.3 + 0.5j, -0.2 + 1.1j, 0.1 + 1.9j]
    ret, _, Ka, bs, nc, rcs = F.nsev_batch(Q, T, 0, None, 1, o, K=np.full(B, 3), Kmax=Kmax, bound_states=G)
    assert ret == 0
    ret1, _, K1, bs1, nc1 = F.nsev(Q[7], T, 0, None, 1, o, K=3, bound_states=G[7], want_contspec=False)
    assert ret1 == 0 and K1 == Ka[7]
    assert np.array_equal(bs[7, :K1], bs1[:K1])
    assert np.array_equal(nc[7, :K1], nc1[:K1])


# ------------------------------------------------------------------ ES4 / TES4
def _match(a, b):
    return np.array([int(np.argmin(np.abs(a - x))) for x in b])


def test_nsev_es4_tes4_vs_reference_runs(F):
    """fnft_nse_discretization_ES4 / _TES4 (fnft__akns_scatter_matrix.c:259-320,464-515,
    fnft__nse_scatter_bound_states.c:124-183,343-470,535-630, preprocessing fnft__nse_discretization.c:609-631):
    continuous spectrum (rho, a, b), Newton bound states with norming constants and residues, Richardson
    extrapolation, against outputs of the unmodified reference (tests/golden/make_golden_es4.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_es4.npz"))
    cases = sorted({tuple(k.split("/")[1:5]) for k in g.files if k.startswith("refrun/slow")})
    assert len(cases) == 14
    for kind, disc, D, kappa in cases:
        key = f"refrun/{kind}/{disc}/{D}/{kappa}"
        q = g[f"refrun/slow/{disc}/{D}/{kappa}/q"]
        gs = g[f"refrun/slow/{disc}/{D}/{kappa}/guesses"]
        o = F.nsev_default_opts()
        o.discretization, o.bound_state_localization, o.discspec_type, o.contspec_type = int(disc), 1, 2, 2
        o.richardson_extrapolation_flag = 1 if kind == "slow_richardson" else 0
        ret, cs, K, bs, nc = F.nsev(q, [-10, 10], 20, [-2, 2.5], int(kappa), o, K=3, bound_states=gs)
        assert ret == 0, key
        ref = g[key + "/cs"]
        for part in range(3):
            assert max(parity_contract(cs[part * 20:(part + 1) * 20], ref[part * 20:(part + 1) * 20])) < 1, key
        rbs, rnc = g[key + "/bs"], g[key + "/nc"]
        assert K == rbs.size, key
        if K:
            idx = _match(bs[:K], rbs)
            assert (np.abs(bs[:K][idx] - rbs) <= 1e-9 * np.abs(rbs)).all(), key
            for part in range(2):
                assert (np.abs(nc[part * K:(part + 1) * K][idx] - rnc[part * K:(part + 1) * K])
                        <= 1e-9 * np.abs(rnc[part * K:(part + 1) * K])).all(), key
    # a batch gives the same values as single calls; oracle at another size
    D = 250
    t = np.linspace(-10, 10, D)
    Q = np.stack([a / np.cosh(t) * np.exp(0.2j * a * t) for a in (0.8, 1.7, 2.6)])
    for disc in (26, 27):
        o = F.nsev_default_opts()
        o.discretization = disc
        o.bound_state_localization = 1
        ret, csb, *_ = F.nsev_batch(Q, [-10, 10], 32, [-2, 2], -1, o)
        assert ret == 0
        for b in range(3):
            r1, cs1, *_ = F.nsev(Q[b], [-10, 10], 32, [-2, 2], -1, o)
            assert r1 == 0 and np.array_equal(cs1, csb[b])
            assert O.misc_rel_err(cs1, O.nsev_contspec_slow(Q[b], [-10, 10], 32, [-2, 2], -1, disc, 0)) < 1e-9


def test_private_scatter_bound_states_es4_tes4(F):
    """fnft__nse_scatter_bound_states with ES4 / TES4 on caller-preprocessed samples (q, q', q'') against the oracle"""
    D, T = 128, (-9.0, 9.0)
    t = np.linspace(T[0], T[1], D)
    q = 1.9 / np.cosh(t) * np.exp(0.3j * t)
    eps_t = (T[1] - T[0]) / (D - 1)
    lam = np.array([0.15 + 1.4j, -0.3 + 0.45j])
    for disc in (26, 27):
        qp = O.preprocess_signal(q, eps_t, 1, disc)
        ret, a, ap, b = F.nse_scatter_bound_states(qp, None, T, lam, disc)
        assert ret == 0
        ra, rap, rb = O.nse_scatter_bound_states(qp, T, lam, 3, None, disc)
        assert np.allclose(a, ra, rtol=1e-10, atol=1e-13) and np.allclose(ap, rap, rtol=1e-10)
        assert np.allclose(b, rb, rtol=1e-9)


# ------------------------------------------------------------------ devices other than 0
def test_config2_shape_on_the_highest_device(F):
    n = F.device_count()
    if n < 2:
        pytest.skip("one GPU visible")
    D = M = 16384
    T, XI = (-32.0, 32.0), (-10.0, 10.0)
    t = np.linspace(T[0], T[1], D)
    Q = np.stack([3.3 / np.cosh(t) * np.exp(-2j * 0.7 * t), 1.2 / np.cosh(t / 1.7) * np.exp(0.5j * np.sin(t))])
    try:
        assert F.set_device(n - 1) == 0
        ret, cs, *_ = F.nsev_batch(Q, T, M, XI, 1)
        assert ret == 0
    finally:
        F.set_device(0)
    ret0, cs0, *_ = F.nsev_batch(Q, T, M, XI, 1)
    assert ret0 == 0 and np.array_equal(cs, cs0)  # same kernels, same bits on every device
    for b in range(2):
        ref = R.nsev(Q[b], T, M, XI, 1, None)[1] if R.available() else O.nsev_contspec(Q[b], T, M, XI, 1)
        assert rel_err(cs[b], ref) < 1e-9


# ------------------------------------------------------------------ several GPUs behind one call
def _fanout_devices(F):
    n = F.device_count()
    return [0, n - 1] if n >= 2 else [0, 0]  # two contexts on one GPU exercise the same host code


def test_fanout_nsev_batch_equals_single_device(F):
    B, D, M, T, XI = 37, 2048, 700, (-14.0, 14.0), (-4.0, 4.0)
    Q = _signals(B, D, T, 3)
    o = F.nsev_default_opts()
    o.contspec_type = F.CSTYPE_BOTH
    ret0, cs0, *_ = F.nsev_batch(Q, T, M, XI, 1, o)
    try:
        assert F.set_devices(_fanout_devices(F) + [0]) == 0  # three shards: 13 + 12 + 12 signals
        got = np.zeros(4, dtype=np.int32)
        assert F.lib().fnft_b200_get_devices(got.ctypes.data_as(C.c_void_p), 4) == 3
        ret1, cs1, *_ = F.nsev_batch(Q, T, M, XI, 1, o)
        # bound states through the fan-out as well (K, bound_states, norming constants are sharded arrays)
        o2 = F.nsev_default_opts()
        o2.bound_state_localization = F.BSLOC_NEWTON
        o2.discspec_type = F.DSTYPE_BOTH
        G = np.tile(np.array([0.2 + 0.6j, -0.1 + 1.4j]), (B, 1))
        r2, _, K2, bs2, nc2, rc2 = F.nsev_batch(Q, T, 0, None, 1, o2, K=np.full(B, 2), Kmax=5, bound_states=G)
    finally:
        F.set_devices([])
    r3, _, K3, bs3, nc3, rc3 = F.nsev_batch(Q, T, 0, None, 1, o2, K=np.full(B, 2), Kmax=5, bound_states=G)
    assert ret0 == 0 and ret1 == 0 and r2 == 0 and r3 == 0
    assert np.array_equal(cs0, cs1)
    assert np.array_equal(K2, K3) and np.array_equal(bs2, bs3) and np.array_equal(nc2, nc3)


def test_two_pipelines_per_device_are_automatic_and_change_nothing(F):
    """Host-buffer batches of at least 2048 signals run as two pipelines (two contexts) on the current device without
    any configuration (fnft_runtime.c: fnftb__fanout_shards); the results are those of the plain path, bit for bit,
    and fnft_b200_get_devices keeps reporting that the caller has not configured devices."""
    B, D, M, T, XI = 2049, 256, 96, (-9.0, 9.0), (-3.0, 3.0)
    Q = _signals(B, D, T, 17)
    ret, cs, *_ = F.nsev_batch(Q, T, M, XI, 1)       # 1025 + 1024 signals on two contexts
    got = np.zeros(4, dtype=np.int32)
    assert F.lib().fnft_b200_get_devices(got.ctypes.data_as(C.c_void_p), 4) == 0
    ret_a, cs_a, *_ = F.nsev_batch(Q[:2047], T, M, XI, 1)  # below the threshold: the caller's own context
    ret_b, cs_b, *_ = F.nsev_batch(Q[2047:], T, M, XI, 1)
    assert ret == 0 and ret_a == 0 and ret_b == 0
    assert np.array_equal(cs[:2047], cs_a) and np.array_equal(cs[2047:], cs_b)
    U = (Q.real * 0.5).astype(np.complex128)
    r1, k1, _ = F.kdvv_batch(U, T, M, (0.1, 3.0))
    r2, k2, _ = F.kdvv_batch(U[:1500], T, M, (0.1, 3.0))
    assert r1 == 0 and r2 == 0 and np.array_equal(k1[:1500], k2)


def test_fanout_kdvv_and_nsep_batch_equal_single_device(F):
    B, D = 11, 1024
    t = np.linspace(-16, 15, D)
    U = np.stack([(0.5 + 0.2 * b) / np.cosh(t - 0.1 * b) ** 2 for b in range(B)]).astype(np.complex128)
    r0, c0, _ = F.kdvv_batch(U, (-16.0, 15.0), 512, (-3.55, 3.95))
    tp = 2 * np.pi / 256 * np.arange(256)
    Qp = np.stack([(0.8 + 0.1 * b) * np.exp(1j * (b % 3) * tp) * (1 + 0.2 * np.cos(tp + b)) for b in range(B)])
    o = F.nsep_default_opts()
    o.localization = 1
    o.filtering = 1
    o.bounding_box[0], o.bounding_box[1], o.bounding_box[2], o.bounding_box[3] = -10, 10, -10, 10
    p0 = F.nsep_batch(Qp, (0.0, 2 * np.pi), 2048, 2048, 1, o)
    try:
        assert F.set_devices(_fanout_devices(F)) == 0
        r1, c1, _ = F.kdvv_batch(U, (-16.0, 15.0), 512, (-3.55, 3.95))
        p1 = F.nsep_batch(Qp, (0.0, 2 * np.pi), 2048, 2048, 1, o)
    finally:
        F.set_devices([])
    assert r0 == 0 and r1 == 0 and np.array_equal(c0, c1)
    assert p0[0] == 0 and p1[0] == 0
    for a, b in zip(p0[1:], p1[1:]):
        assert np.array_equal(a, b)


def test_fanout_error_codes_come_back_per_signal(F):
    """a(xi) = 0 cannot be provoked portably; an invalid option must come back through the shards"""
    Q = _signals(8, 256, (-8.0, 8.0), 9)
    o = F.nsev_default_opts()
    o.discretization = 9999
    try:
        F.set_devices(_fanout_devices(F))
        ret, *_ = F.nsev_batch(Q, (-8.0, 8.0), 64, (-2.0, 2.0), 1, o)
    finally:
        F.set_devices([])
    assert ret != 0


def test_nsep_batch_with_device_pointers_is_refused_clearly(F):
    L = F.lib()
    Q = np.ones((2, 64), dtype=np.complex128)
    try:
        L.fnft_b200_set_device_pointers(1)
        ret = F.nsep_batch(Q, (0.0, 2 * np.pi), 64, 64, 1)[0]
    finally:
        L.fnft_b200_set_device_pointers(0)
    assert ret == 6  # FNFT_EC_NOT_YET_IMPLEMENTED


def test_fp64_probe_reports_a_plausible_peak(F):
    tf = float(F.lib().fnft_b200_probe_fp64_tflops())
    assert 20.0 < tf < 60.0, tf   # B200: 37-40 TFLOP/s nominal


# ------------------------------------------------------------------ bench.py's parity gate
def test_bench_parity_gate_trips_on_a_corrupted_output():
    """bench.py must exit non-zero when an output differs from the reference (--corrupt adds 1.0 to one
    value) and zero otherwise; small batch, no extras, reference sample of a few signals."""
    if not R.available():
        pytest.skip("oracle/_ref not present on this box")
    import json
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--batch", "128",
           "--no-extras"]
    env = dict(os.environ)
    good = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert good.returncode == 0, good.stderr[-2000:]
    line = json.loads(good.stdout.strip().splitlines()[-1])
    assert line["parity"]["ok"] and line["parity"]["max"] <= 1e-9 and line["parity"]["signals"] >= 16
    bad = subprocess.run(cmd + ["--corrupt"], capture_output=True, text=True, env=env, timeout=900)
    assert bad.returncode == 1, (bad.returncode, bad.stderr[-2000:])
    line = json.loads(bad.stdout.strip().splitlines()[-1])
    assert not line["parity"]["ok"] and line["parity"]["checks"]["config2_device"]["max"] > 1e-9


def test_default_options_root_finder_cta_sizes_agree(F):
    """The compacting root finder (poly_roots.cuh, k_roots_aberth_c) picks its CTA size by the number of polynomials in
    flight: 256 threads from 296 polynomials, 512 from 74, one CTA with as many threads as roots for fewer.  The three
    schedules update the roots in different orders; after the Newton refinement of SUBSAMPLE_AND_REFINE the default
    fnft_nsev must return the same bound states whichever one ran."""
    D, T = 2048, (-14.0, 14.0)   # Dsub = 499: a(z) of degree 998, CTAs of 256 / 512 / 1024 threads
    B = 320
    Q = _signals(B, D, T, 77)
    o = F.nsev_default_opts()
    Kmax = 16
    K0, G0 = np.zeros(B), np.zeros((B, Kmax), dtype=np.complex128)
    ret_a, _, Ka, bsa, _, rca = F.nsev_batch(Q, T, 0, None, 1, o, K=K0, Kmax=Kmax, bound_states=G0)       # 256 threads
    nb = 96
    ret_b, _, Kb, bsb, _, rcb = F.nsev_batch(Q[:nb], T, 0, None, 1, o, K=K0[:nb], Kmax=Kmax, bound_states=G0[:nb])  # 512
    assert ret_a == 0 and ret_b == 0 and (rca == 0).all() and (rcb == 0).all()
    assert (Ka[:nb] == Kb).all() and Ka.sum() >= B          # sech pulses of amplitude 0.6 ... 3: about two each
    for i in range(nb):
        k = int(Ka[i])
        assert np.abs(np.sort_complex(bsa[i, :k]) - np.sort_complex(bsb[i, :k])).max() <= 1e-10 if k else True
    for i in (0, 5, 17):                                      # single calls: one CTA per polynomial, 1024 threads
        r1, _, K1, bs1, _ = F.nsev(Q[i], T, 0, None, 1, o, K=Kmax, want_contspec=False)
        assert r1 == 0 and K1 == int(Ka[i])
        if K1:
            assert np.abs(np.sort_complex(bs1[:K1]) - np.sort_complex(bsa[i, :K1])).max() <= 1e-10

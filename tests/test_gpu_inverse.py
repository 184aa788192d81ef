"""Inverse NFT on the GPU: fnft_nsev_inverse, fnft__nse_finvscatter, fnft__poly_specfact through the
C-ABI against the UNMODIFIED reference (oracle/_ref) on the same inputs, and the reference's own round-trip property
(fnft__nse_finvscatter_test.inc: forward scattering followed by inverse scattering returns the samples to 1262 eps
at D = 16384).  The reference's inverse test programs themselves run in tests/test_reference_programs.py."""
import numpy as np
import pytest

from common import ensure_lib, rel_err
import inverse_bindings as IB
from oracle import fnft_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps
D2A, MODAL = 4, 0   # fnft_nse_discretization_2SPLIT2A / _2SPLIT2_MODAL


@pytest.fixture(scope="module")
def F():
    ensure_lib()
    import fnft_b200 as F_
    F_.quiet(True)
    return F_


@pytest.fixture(scope="module")
def RL():
    from oracle import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref/libfnft_ref.so not built")
    ref_lib.lib().fnft_errwarn_setprintf(None)
    return ref_lib.lib()


@pytest.mark.parametrize("disc", [D2A, MODAL])
@pytest.mark.parametrize("kappa", [+1, -1])
@pytest.mark.parametrize("D", [2, 8, 512, 1024, 16384])
def test_finvscatter_round_trip_like_the_reference_test(F, RL, D, kappa, disc):
    # test/fnft__nse_finvscatter/fnft__nse_finvscatter_test.inc:29-72, error bound of the Kiss FFT build
    i = np.arange(D)
    q_exact = ((i + 1) / (D + 1) / D) * np.exp(1j * i / D)
    eps_t = 0.12
    ret, tm, deg, W = F.nse_fscatter(q_exact, eps_t, kappa, disc, normalize=False)
    assert ret == 0 and deg == D
    ret, q = IB.nse_finvscatter(F.lib(), tm, eps_t, kappa, disc)
    assert ret == 0
    # the reference's bounds for its Kiss FFT build (test/fnft__nse_finvscatter/*.c:25-29); the defocusing 2SPLIT2A
    # bound is larger because the base case maps |Q| back with atan also when |Q| = tanh(eps |q|) (:178)
    bound = {(+1, D2A): 1262, (+1, MODAL): 1150, (-1, D2A): 81065, (-1, MODAL): 1253}[(kappa, disc)]
    r2, q_ref = IB.nse_finvscatter(RL, tm, eps_t, kappa, disc)
    assert r2 == 0
    if (kappa, disc) == (-1, D2A):
        # atan instead of atanh: a relative error of (2/3) (eps |q|)^2 that is the reference's, reproduced as is;
        # it only drops below the reference's bound at the D = 16384 of the reference's test (|q| ~ 1/D)
        assert rel_err(q, q_ref) < 1262 * EPS
        if D == 16384:
            assert rel_err(q, q_exact) < bound * EPS
    else:
        assert rel_err(q, q_exact) < bound * EPS
        assert rel_err(q, q_ref) < 2 * bound * EPS


@pytest.mark.parametrize("disc", [D2A, MODAL])
@pytest.mark.parametrize("D", [256, 4096])
def test_finvscatter_vs_reference_on_a_pulse(F, RL, D, disc):
    # a sech pulse with a chirp, amplitude of order one: the same transfer matrix into both implementations
    t = np.linspace(-8, 8, D)
    q0 = 1.3 / np.cosh(t) * np.exp(0.7j * t * t)
    eps_t = t[1] - t[0]
    for kappa in (+1, -1):
        qq = q0 if kappa > 0 else 0.3 * q0   # |eps q| < 1 is required in the defocusing modal case
        ret, tm, deg, W = F.nse_fscatter(qq, eps_t, kappa, disc, normalize=False)
        assert ret == 0
        r1, q1 = IB.nse_finvscatter(F.lib(), tm, eps_t, kappa, disc)
        r2, q2 = IB.nse_finvscatter(RL, tm, eps_t, kappa, disc)
        assert r1 == 0 and r2 == 0
        # both recover the signal; the comparison is with the signal (each implementation's own rounding errors
        # are amplified by the layer peeling in the same way, not to the same values)
        e1, e2 = rel_err(q1, qq), rel_err(q2, qq)
        assert e1 < max(10 * e2, 1e-11), (kappa, e1, e2)


def test_finvscatter_reports_a_sample_outside_the_unit_disc(F):
    # defocusing case, |Q| >= 1: src/private/fnft__nse_finvscatter.c:172-176 returns an error
    D = 16
    tm = np.zeros((4, D + 1), dtype=np.complex128)
    tm[0, D] = 1.0
    tm[2, D] = 2.0   # Q = conj(T21 / T11) = 2
    tm[3, 0] = 1.0
    ret, q = IB.nse_finvscatter(F.lib(), tm, 0.1, -1, D2A)
    assert ret != 0


@pytest.mark.parametrize("kappa", [+1, -1, 0])
@pytest.mark.parametrize("deg", [7, 255, 1000])
def test_poly_specfact_vs_reference(F, RL, deg, kappa):
    rng = np.random.default_rng(deg + 3 * kappa)
    p = (rng.standard_normal(deg + 1) + 1j * rng.standard_normal(deg + 1)) * np.exp(-0.05 * np.arange(deg + 1))
    p *= 0.4 / np.abs(p).sum()    # |P| < 1 on the unit circle (kappa = +1)
    if kappa == 0:
        p[0] += 1.0               # no zeros on the unit circle
    for ovs in (4, 8):
        r1, a1 = IB.poly_specfact(F.lib(), p, ovs, kappa)
        r2, a2 = IB.poly_specfact(RL, p, ovs, kappa)
        assert r1 == 0 and r2 == 0
        # the oracle restatement uses an exact FFT at the reference's length; where that length is no power of two
        # (deg 1000: 4050, 8100) the REFERENCE is 2e-11 ... 3e-10 away from it (Kiss FFT radix-3/5 butterflies), we
        # are not
        M = O.next_fast_size((deg + 1) * ovs)
        pow2 = (M & (M - 1)) == 0
        assert rel_err(a1, O.poly_specfact(p, ovs, kappa)) < 1e-13
        assert rel_err(a1, a2) < (1e-13 if pow2 else 1e-9), (ovs, rel_err(a1, a2))


def _solitons(K, seed):
    rng = np.random.default_rng(seed)
    bs = rng.uniform(-1.5, 1.5, K) + 1j * rng.uniform(0.3, 2.0, K)
    nc = np.exp(rng.uniform(-1, 1, K) + 1j * rng.uniform(0, 2 * np.pi, K))
    return bs, nc


@pytest.mark.parametrize("K", [1, 3, 8, 20])
@pytest.mark.parametrize("dstype", [0, 1])
def test_pure_multisoliton_vs_reference(F, RL, K, dstype):
    # src/fnft_nsev_inverse.c:803-846 (no continuous spectrum); T not symmetric so that both branches run
    bs, nc = _solitons(K, K)
    D, T = 1024, (-9.0, 11.0)
    for method in (0, 3):   # DEFAULT (closed-form recursion) and USE_SEED_POTENTIAL_INSTEAD (zero seed + Darboux)
        o1, o2 = IB.default_opts(F.lib()), IB.default_opts(RL)
        for o in (o1, o2):
            o.discspec_type = dstype
            o.contspec_inversion_method = method
        r1, q1, _ = IB.nsev_inverse(F.lib(), None, None, bs, nc, D, T, +1, o1)
        r2, q2, _ = IB.nsev_inverse(RL, None, None, bs, nc, D, T, +1, o2)
        assert r1 == 0 and r2 == 0
        print("multisoliton K", K, "dstype", dstype, "method", method, "rel_err", rel_err(q1, q2))
        # 20 solitons from residues: the conversion to norming constants multiplies 19 ratios (1.1e-10 measured)
        assert rel_err(q1, q2) < (1e-9 if K >= 20 else 1e-11), (method, rel_err(q1, q2))


def test_config3_generator_batch_equals_reference(F, RL):
    # what bench.py's config 3 asks the reference for (8 solitons, D = 4096, norming constants): the batched GPU
    # entry point returns the same signals
    import bench as BM
    lams, bn, _ = BM.config3_params()
    n = 16
    o = IB.default_opts(F.lib())
    ret, q, rcs = IB.nsev_inverse_batch(F.lib(), None, None, lams[:n], bn[:n], BM.C3["D"], BM.C3["T"], +1, o)
    assert ret == 0 and (rcs == 0).all()
    for i in range(0, n, 5):
        qr = BM._soliton_worker((lams[i], bn[i], BM.C3["D"], BM.C3["T"]))
        assert rel_err(q[i], qr) < 1e-11


def _contspec_case(L, D, M, T, kappa, disc, cstype):
    """a continuous spectrum on the grid of fnft_nsev_inverse_XI: rho, b(xi) or B(tau) of a small sech-like pulse"""
    ret, XI = IB.inverse_XI(L, D, T, M, disc)
    assert ret == 0
    xi = np.linspace(XI[0], XI[1], M)
    if cstype == 2:   # B(tau) samples on the time grid (M == D)
        tau = np.linspace(T[0], T[1], D)
        return XI, 0.4 / np.cosh(2 * tau) * np.exp(0.5j * tau)
    amp = 0.5 if cstype == 0 else 0.35
    return XI, amp / np.cosh(np.pi * xi / 2) * np.exp(0.3j * xi)


@pytest.mark.parametrize("disc", [D2A, MODAL])
@pytest.mark.parametrize("cstype,method,kappa", [(0, 0, +1), (0, 1, -1), (0, 2, -1), (1, 0, +1), (1, 0, -1),
                                                 (2, 0, +1), (2, 0, -1)])
def test_inverse_from_continuous_spectrum_vs_reference(F, RL, cstype, method, kappa, disc):
    D = 512
    M = D if (cstype == 2 or method == 2) else 2 * D
    T = (-8.0, 8.0)
    XI, cs = _contspec_case(F.lib(), D, M, T, kappa, disc, cstype)
    o1, o2 = IB.default_opts(F.lib()), IB.default_opts(RL)
    for o in (o1, o2):
        o.discretization = disc
        o.contspec_type = cstype
        o.contspec_inversion_method = method
    r1, q1, c1 = IB.nsev_inverse(F.lib(), cs, XI, None, None, D, T, kappa, o1)
    r2, q2, c2 = IB.nsev_inverse(RL, cs, XI, None, None, D, T, kappa, o2)
    assert r1 == 0 and r2 == 0
    # the in-place modification of contspec is part of the interface (phases up to ~800 rad: the last bit of xi shows)
    assert rel_err(c1, c2) < 1e-12
    print("contspec", cstype, method, kappa, disc, "rel_err q", rel_err(q1, q2))
    assert rel_err(q1, q2) < 1e-9, rel_err(q1, q2)


@pytest.mark.parametrize("cstype,dstype,method", [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (2, 0, 0), (0, 0, 3)])
def test_inverse_with_both_spectra_vs_reference(F, RL, cstype, dstype, method):
    D = 512
    M = D if cstype == 2 else 2 * D
    T = (-8.0, 8.0)
    XI, cs = _contspec_case(F.lib(), D, M, T, +1, D2A, cstype)
    bs, nc = _solitons(3, 11)
    o1, o2 = IB.default_opts(F.lib()), IB.default_opts(RL)
    for o in (o1, o2):
        o.contspec_type = cstype
        o.discspec_type = dstype
        o.contspec_inversion_method = method
    r1, q1, _ = IB.nsev_inverse(F.lib(), cs, XI, bs, nc, D, T, +1, o1)
    r2, q2, _ = IB.nsev_inverse(RL, cs, XI, bs, nc, D, T, +1, o2)
    assert r1 == r2
    if r1 == 0:
        assert rel_err(q1, q2) < 1e-8, rel_err(q1, q2)


def test_inverse_argument_checks_like_the_reference(F, RL):
    # src/fnft_nsev_inverse.c:134-172: same return codes for the same bad arguments
    bs, nc = _solitons(2, 5)
    cs = np.ones(64, dtype=np.complex128)
    XI = np.array([-1.0, 1.0])
    cases = [
        dict(contspec=cs[:63], XI=XI, bs=None, nc=None, D=32, T=(-1, 1), kappa=1),       # odd M
        dict(contspec=cs[:16], XI=XI, bs=None, nc=None, D=32, T=(-1, 1), kappa=1),       # M < D
        dict(contspec=cs, XI=XI, bs=None, nc=None, D=48, T=(-1, 1), kappa=1),            # D not a power of two
        dict(contspec=None, XI=None, bs=bs, nc=nc, D=32, T=(1, -1), kappa=1),            # T
        dict(contspec=None, XI=None, bs=bs, nc=nc, D=32, T=(-1, 1), kappa=-1),           # solitons, defocusing
        dict(contspec=None, XI=None, bs=np.conj(bs), nc=nc, D=32, T=(-1, 1), kappa=1),   # lower half plane
        dict(contspec=None, XI=None, bs=None, nc=None, D=32, T=(-1, 1), kappa=1),        # nothing given
        dict(contspec=cs, XI=None, bs=None, nc=None, D=32, T=(-1, 1), kappa=1),          # XI missing
    ]
    F.lib().fnft_errwarn_setprintf(None)
    for c in cases:
        o1, o2 = IB.default_opts(F.lib()), IB.default_opts(RL)
        r1, _, _ = IB.nsev_inverse(F.lib(), c["contspec"], c["XI"], c["bs"], c["nc"], c["D"], c["T"], c["kappa"], o1)
        r2, _, _ = IB.nsev_inverse(RL, c["contspec"], c["XI"], c["bs"], c["nc"], c["D"], c["T"], c["kappa"], o2)
        assert r1 == r2 and r1 != 0, (c, r1, r2)
    o1, o2 = IB.default_opts(F.lib()), IB.default_opts(RL)
    o1.discretization = o2.discretization = 11     # 2SPLIT4B is not invertible (:164-166)
    r1, _, _ = IB.nsev_inverse(F.lib(), None, None, bs, nc, 32, (-1, 1), 1, o1)
    r2, _, _ = IB.nsev_inverse(RL, None, None, bs, nc, 32, (-1, 1), 1, o2)
    assert r1 == r2 and r1 != 0


def test_inverse_vs_oracle_restatements(F):
    # the numpy restatements of oracle/fnft_oracle.py (pinned to the reference in tests/test_oracle.py)
    D = 64
    t = np.linspace(-4, 4, D)
    q0 = 0.9 / np.cosh(t) * np.exp(0.4j * t)
    eps_t = t[1] - t[0]
    for disc in (D2A, MODAL):
        for kappa in (+1, -1):
            qq = q0 if kappa > 0 else 0.5 * q0
            ret, tm, deg, W = F.nse_fscatter(qq, eps_t, kappa, disc, normalize=False)
            assert ret == 0
            ret, q = IB.nse_finvscatter(F.lib(), tm, eps_t, kappa, disc)
            assert ret == 0 and rel_err(q, O.nse_finvscatter(tm, eps_t, kappa, disc)) < 1e-11
    bs, nc = _solitons(5, 3)
    for dstype in (0, 1):
        o = IB.default_opts(F.lib())
        o.discspec_type = dstype
        ret, q, _ = IB.nsev_inverse(F.lib(), None, None, bs, nc, 256, (-7.0, 9.0), +1, o)
        assert ret == 0
        assert rel_err(q, O.nsev_inverse_pure_solitons(bs, nc, 256, (-7.0, 9.0), residues=bool(dstype))) < 1e-12

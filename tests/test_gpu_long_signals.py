"""GPU tests (-m gpu) of signals LONGER than one product tree can hold (transfer matrix of degree > 2^18):
the reference multiplies polynomials of any length (src/private/fnft__poly_fmult.c:404-445); the drop-in cuts
the signal into pieces, runs each through the normal path and chains the pieces' scattering coefficients on the
xi grid (fnft_nsev.c: nsev_contspec_segmented).  FNFT_B200_TREE_MAX_SAMPLES lowers the limit so that the
segmented path can be compared with the direct one at sizes both handle; the knob is read once per process,
hence the subprocesses."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import parity_contract, rel_err
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
    return np.stack([rng.uniform(0.6, 2.2) / np.cosh(t / rng.uniform(0.6, 1.6) - rng.uniform(-1, 1)) *
                     np.exp(1j * rng.uniform(-2, 2) * t + 1j * rng.uniform(0, 6.28)) +
                     0.02 * (rng.standard_normal(D) + 1j * rng.standard_normal(D)) for _ in range(B)])


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
import fnft_b200 as F
F.lib().fnft_errwarn_setprintf(None)
d = np.load(%(inp)r)
Q, T, XI, M = d["Q"], d["T"], d["XI"], int(d["M"])
out = {}
for kappa in (+1, -1):
    for cst in (0, 1, 2):
        o = F.nsev_default_opts()
        o.contspec_type = cst
        ret, cs, _, _, _, rcs = F.nsev_batch(Q, T, M, XI, kappa, o)
        assert ret == 0 and (rcs == 0).all(), (ret, rcs)
        out["b_%%d_%%d" %% (kappa, cst)] = cs
        ret, cs1, _, _, _ = F.nsev(Q[1], T, M, XI, kappa, o)
        assert ret == 0
        out["s_%%d_%%d" %% (kappa, cst)] = cs1
U = (Q.real * 0.8 + 0.3 * np.abs(Q)).astype(np.complex128)
for disc in (%(kdv_discs)s):
    o = F.kdvv_default_opts()
    o.discretization = disc
    ret, cs, rcs = F.kdvv_batch(U, T, M, XI, o)
    assert ret == 0 and (rcs == 0).all(), (ret, rcs)
    out["kdv_b_%%d" %% disc] = cs
    ret, cs1 = F.kdvv(U[2], T, M, XI, o)
    assert ret == 0
    out["kdv_s_%%d" %% disc] = cs1
np.savez(%(outp)r, **out)
"""
KDV_DISCS = "2, 9, 19"  # fnft_kdv_discretization_2SPLIT2A (sqrt(z) correction), 2SPLIT4B, 4SPLIT4B


def _run_child(tmp_path, tag, Q, T, XI, M, limit):
    inp, outp = str(tmp_path / (tag + "_in.npz")), str(tmp_path / (tag + "_out.npz"))
    np.savez(inp, Q=Q, T=np.array(T), XI=np.array(XI), M=M)
    env = dict(os.environ)
    env.pop("FNFT_B200_TREE_MAX_SAMPLES", None)
    if limit:
        env["FNFT_B200_TREE_MAX_SAMPLES"] = str(limit)
    r = subprocess.run([sys.executable, "-c", _CHILD % {"root": ROOT, "inp": inp, "outp": outp, "kdv_discs": KDV_DISCS}], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(outp)


@pytest.mark.parametrize("D,limit", [(1000, 256), (4099, 1024)])
def test_segmented_contspec_equals_direct_path_and_reference(F, tmp_path, D, limit):
    """4 and 5 ragged pieces against the one-tree result of the same library and against the reference, for the
    three continuous-spectrum types and both signs of kappa, batched and single."""
    T, XI, M = (-9.0, 11.0), (-4.0, 3.0), 257
    Q = _signals(3, D, T, 21)
    direct = _run_child(tmp_path, "direct", Q, T, XI, M, 0)
    seg = _run_child(tmp_path, "seg", Q, T, XI, M, limit)
    for key in direct.files:
        a, b = np.asarray(direct[key]), np.asarray(seg[key])
        assert a.shape == b.shape
        for j in range(a.shape[0] if a.ndim == 2 else 1):
            x, y = (a[j], b[j]) if a.ndim == 2 else (a, b)
            n = x.shape[0] // M
            for part in range(n):  # rho, a, b separately: their magnitudes differ by orders
                assert max(parity_contract(y[part * M:(part + 1) * M], x[part * M:(part + 1) * M])) < 1, (key, j, part)
    if R.available():
        for kappa in (+1, -1):
            o = R.nsev_default_opts()
            o.contspec_type = 2
            ret, ref, _, _, _ = R.nsev(Q[1], np.array(T), M, np.array(XI), kappa, o)
            assert ret == 0
            ours = seg["s_%d_2" % kappa]
            for part in range(3):
                assert max(parity_contract(ours[part * M:(part + 1) * M], ref[part * M:(part + 1) * M])) < 1


def test_signal_longer_than_one_tree_against_the_reference(F):
    """D = 150 001 > 131 072 (2SPLIT4B: degree 300 002 > 2^18): two pieces of 75 001 and 75 000 samples."""
    D, T, XI, M = 150001, (-40.0, 40.0), (-6.0, 6.0), 96
    q = _signals(1, D, T, 33)[0]
    o = F.nsev_default_opts()
    o.contspec_type = 2
    ret, cs, _, _, _ = F.nsev(q, T, M, XI, 1, o)
    assert ret == 0
    assert np.isfinite(cs.view(np.float64)).all()
    rho, a, b = cs[:M], cs[M:2 * M], cs[2 * M:]
    # |a|^2 + kappa |b|^2 = 1 on the real axis (unimodular transfer matrix), rho = b / a
    assert np.abs(np.abs(a) ** 2 + np.abs(b) ** 2 - 1).max() < 1e-9
    assert rel_err(rho, b / a) < 1e-13
    # Independent truth: the product of the per-sample leaf matrices (oracle.akns_leaves, src/private/
    # fnft__akns_fscatter.c:402-433) applied to (1, 0) sample by sample in long double at every fourth xi --
    # no FFT products, no chirp-z.
    idx = np.arange(0, M, 4)
    eps_t = (T[1] - T[0]) / (D - 1)
    xi = (XI[0] + (XI[1] - XI[0]) / (M - 1) * idx).astype(np.longdouble)
    z = np.exp(1j * xi * np.longdouble(eps_t))  # lambda_to_z: exp(2i xi eps_t / deg), deg = 2
    z2 = z * z
    P = O.akns_leaves(q, -np.conj(q), eps_t, O.AKNS_2SPLIT4B).astype(np.clongdouble)  # [4][D][3], matrix k = sample D-1-k
    v1, v2 = np.ones(len(idx), dtype=np.clongdouble), np.zeros(len(idx), dtype=np.clongdouble)
    for k in range(D - 1, -1, -1):
        m11 = P[0, k, 0] * z2 + P[0, k, 1] * z + P[0, k, 2]
        m12 = P[1, k, 0] * z2 + P[1, k, 1] * z + P[1, k, 2]
        m21 = P[2, k, 0] * z2 + P[2, k, 1] * z + P[2, k, 2]
        m22 = P[3, k, 0] * z2 + P[3, k, 1] * z + P[3, k, 2]
        v1, v2 = m11 * v1 + m12 * v2, m21 * v1 + m22 * v2
    ph_rho = np.longdouble(-2.0 * (T[1] + 0.5 * eps_t))  # src/private/fnft__nse_discretization.c:240-256
    truth = (v2 / v1 * np.exp(1j * xi * ph_rho)).astype(np.complex128)
    # the pieces are evaluated exactly on the unit circle, so the chained result is much closer to the truth than
    # 1e-9 (measured 3e-13; scripts/long_accuracy.py: the same up to D = 1 000 003)
    assert max(parity_contract(rho[idx], truth, tol=1e-11)) < 1
    if R.available():
        # the reference evaluates at its rounded A V^-m, a few 1e-17 off the unit circle: |z|^deg - 1 ~ m deg 1e-17 is its
        # distance from the truth here (0.3e-9), and therefore ours from the reference
        ro = R.nsev_default_opts()
        ro.contspec_type = 2
        rret, ref, _, _, _ = R.nsev(q, np.array(T), M, np.array(XI), 1, ro)
        assert rret == 0
        for part in range(3):
            assert max(parity_contract(cs[part * M:(part + 1) * M], ref[part * M:(part + 1) * M])) < 1, part
        print("long signal: ours vs truth %.2e, reference vs truth %.2e (x 1e-9)" %
              (max(parity_contract(rho[idx], truth)), max(parity_contract(ref[:M][idx], truth))))


@pytest.mark.parametrize("disc,D", [(16, 3000), (18, 12000)])
def test_high_degree_schemes_beyond_one_tree(F, disc, D):
    """2SPLIT7A (degree 105 per sample, padded to 128: one tree holds 2048 samples) at D = 3000 and 2SPLIT8A (degree 24,
    8192 samples) at D = 12000 against the reference."""
    if not R.available():
        pytest.skip("oracle/_ref not present on this box")
    T, XI, M = (-10.0, 10.0), (-3.0, 3.0), 48
    q = _signals(1, D, T, 77)[0]
    o = F.nsev_default_opts()
    o.discretization = disc
    o.contspec_type = 2
    ret, cs, _, _, _ = F.nsev(q, T, M, XI, 1, o)
    assert ret == 0
    ro = R.nsev_default_opts()
    ro.discretization = disc
    ro.contspec_type = 2
    rret, ref, _, _, _ = R.nsev(q, np.array(T), M, np.array(XI), 1, ro)
    assert rret == 0
    for part in range(3):
        assert max(parity_contract(cs[part * M:(part + 1) * M], ref[part * M:(part + 1) * M])) < 1, part


def test_kdvv_longer_than_one_tree_against_the_reference(F):
    """fnft_kdvv, D = 150 001: two pieces, general 2x2 chaining of the raw transfer-matrix values."""
    D, T, XI, M = 150001, (-40.0, 40.0), (0.2, 5.0), 64
    t = np.linspace(T[0], T[1], D)
    rng = np.random.default_rng(5)
    u = (1.1 / np.cosh(t - 0.3) ** 2 + 0.4 / np.cosh(0.7 * (t + 9.0)) ** 2 + 0.01 * rng.standard_normal(D)).astype(np.complex128)
    ret, cs = F.kdvv(u, T, M, XI, None)
    assert ret == 0 and np.isfinite(cs.view(np.float64)).all()
    assert np.abs(cs).max() > 1e-3
    if R.available():
        rret, ref = R.kdvv(u, np.array(T), M, np.array(XI))
        assert rret == 0
        assert max(parity_contract(cs, ref)) < 1


def test_long_signal_bound_states_with_newton(F):
    """Newton refinement never touches the polynomial, so it works at any length next to the segmented
    continuous spectrum; the other localizations are refused with a clear error code."""
    D, T, XI, M = 140000, (-30.0, 30.0), (-2.0, 2.0), 32
    t = np.linspace(T[0], T[1], D)
    q = 2.3 / np.cosh(t)
    o = F.nsev_default_opts()
    o.bound_state_localization = F.BSLOC_NEWTON
    ret, cs, K, bs, nc = F.nsev(q, T, M, XI, 1, o, K=2, bound_states=np.array([0.4j, 1.7j]))
    assert ret == 0 and K == 2
    assert np.abs(np.sort(bs.imag) - np.array([0.8, 1.8])).max() < 1e-6 and np.abs(bs.real).max() < 1e-9
    # the default options (SUBSAMPLE_AND_REFINE: fast eigenvalues of the subsampled signal, Newton on the full one)
    od = F.nsev_default_opts()
    ret, cs2, K2, bs2, nc2 = F.nsev(q, T, M, XI, 1, od, K=64)
    assert ret == 0 and K2 == 2, (ret, K2)
    assert np.abs(np.sort(bs2.imag) - np.array([0.8, 1.8])).max() < 1e-6
    assert max(parity_contract(cs2, cs)) < 1
    o.bound_state_localization = F.BSLOC_FAST_EIGENVALUE
    ret, _, _, _, _ = F.nsev(q, T, M, XI, 1, o, K=16)
    assert ret != 0

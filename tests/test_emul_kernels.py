"""CPU checks of the CUDA block programs through the host emulation build
(tests/emul/emul_lib.cpp, -DFNFTB_EMUL): the same kernel source with the thread loop made
explicit.  Verifies the index arithmetic of the FFT passes, the product tree (direct,
in-shared-memory and row-split paths, padding, wrap correction, lazy normalisation) and
the four-step chirp-z against the oracle.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from common import (AKNS_TEST_BOUND, AKNS_TEST_SCHEMES, akns_fscatter_test_input, ensure_emul, eval_tm,
                    rel_err)
from oracle import fnft_oracle as O

dp = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def E():
    L = C.CDLL(ensure_emul())
    L.emul_fft.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, ip]
    L.emul_fscatter.argtypes = [dp, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_double, C.c_int, dp, ip, C.c_int, C.c_int]
    L.emul_fmult2x2.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, ip, C.c_int, C.c_int]
    L.emul_chirpz.argtypes = ([dp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
                              + [C.c_double] * 4 + [C.c_int, C.c_int, dp, C.c_long, C.c_void_p]
                              + [C.c_double] * 7 + [C.c_void_p, C.c_int])
    return L


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_shared_memory_fft_passes(E, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
    d = x.copy()
    perm = np.zeros(n, dtype=np.int32)
    E.emul_fft(d, 3, n, -1, 64, perm)
    ref = np.fft.fft(x, axis=1)
    assert sorted(perm) == list(range(n))
    assert np.abs(d - ref[:, perm]).max() <= 1e-14 * np.abs(ref).max() * max(1, np.log2(n))
    E.emul_fft(d, 3, n, +1, 64, perm)           # inverse consumes the permuted order
    assert np.abs(d / n - x).max() <= 1e-14 * max(1, np.log2(n))


NSE2AKNS = {11: 10, 4: 3, 0: 0, 2: 1, 3: 2, 5: 4, 6: 5}


@pytest.mark.parametrize("disc", [11, 4, 0, 5])
@pytest.mark.parametrize("D", [2, 3, 37, 64, 100, 300])
@pytest.mark.parametrize("use_direct,smem_n", [(1, 1024), (0, 1024), (0, 16), (1, 32)])
def test_tree_kernels_vs_oracle(E, disc, D, use_direct, smem_n):
    rng = np.random.default_rng(D)
    t = np.linspace(-5, 5, D)
    eps_t = 10.0 / max(D - 1, 1)
    q = np.stack([2.3 / np.cosh(t) * np.exp(0.7j * t),
                  (rng.standard_normal(D) + 1j * rng.standard_normal(D)) * 0.5])
    ak = NSE2AKNS[disc]
    d0 = O.akns_degree(ak)
    # the modal scheme rejects kappa = -1 with eps_t*|q| >= 1 (akns_fscatter.c:122-126)
    for kappa in ((1,) if disc == 0 else (1, -1)):
        tm = np.zeros((2, 4, d0 * D + 1), dtype=np.complex128)
        W = np.zeros(2, dtype=np.int32)
        rc = E.emul_fscatter(np.ascontiguousarray(q), None, 2, D, d0, 0, kappa, ak, eps_t, 1, tm, W,
                             use_direct, smem_n)
        assert rc == 0
        for s in range(2):
            tmo, dego, Wo = O.nse_fscatter(q[s], eps_t, kappa, disc)
            for e in range(4):
                if np.abs(tmo[e]).sum() > 0:
                    assert rel_err(tm[s, e] * 2.0 ** W[s], tmo[e] * 2.0 ** Wo) < 1e-12


@pytest.mark.parametrize("deg,n", [(1, 4), (1, 5), (3, 5), (5, 9), (7, 33)])
@pytest.mark.parametrize("use_direct,smem_n", [(1, 1024), (0, 16)])
def test_general_fmult2x2_vs_oracle(E, deg, n, use_direct, smem_n):
    rng = np.random.default_rng(deg * 100 + n)
    p = rng.standard_normal((4, n, deg + 1)) + 1j * rng.standard_normal((4, n, deg + 1))
    for normalize in (0, 1):
        tm = np.zeros((4, deg * n + 1), dtype=np.complex128)
        W = np.zeros(1, dtype=np.int32)
        assert E.emul_fmult2x2(np.ascontiguousarray(p), n, deg, normalize, tm, W, use_direct, smem_n) == 0
        res, dego, Wo = O.poly_fmult2x2(p, bool(normalize))
        assert dego == deg * n
        assert rel_err((tm * 2.0 ** W[0]).reshape(-1), (res * 2.0 ** Wo).reshape(-1)) < 1e-12


@pytest.mark.parametrize("deg,M,row_n", [(3, 3, 4096), (3, 6, 4096), (100, 57, 4096), (100, 57, 16),
                                         (1000, 1500, 64), (2048, 700, 256),
                                         # short polynomial, long transform: direct column evaluation
                                         (20, 1000, 16), (40, 3000, 16), (63, 900, 16)])
def test_chirpz_four_step_vs_oracle(E, deg, M, row_n):
    rng = np.random.default_rng(deg + M)
    p = rng.standard_normal(deg + 1) + 1j * rng.standard_normal(deg + 1)
    A, W = np.exp(-1.3j), np.exp(0.002j)
    out = np.zeros(M, dtype=np.complex128)
    rc = E.emul_chirpz(p, deg + 1, 0, 0, 1, deg, 1, M, 0.0, np.angle(W), 0.0, np.angle(A), 0, 0, out, M,
                       None, 0, 0, 0, 0, 0, 0, 0, None, row_n)
    assert rc == 0
    z = 1.0 / (A * W ** (-np.arange(M)))
    exact = np.polyval(p, z)
    assert rel_err(out, exact) < 1e-9


@pytest.mark.parametrize("D", [2, 5, 64, 130, 300])
@pytest.mark.parametrize("use_direct,smem_n", [(1, 1024), (0, 16), (1, 32)])
def test_tree_kernels_general_mode_vs_oracle(E, D, use_direct, smem_n):
    """KdV (r = -1) and explicit-r inputs take the general 4-entry path (no symmetry)."""
    rng = np.random.default_rng(1000 + D)
    t = np.linspace(-6, 6, D)
    eps_t = 12.0 / max(D - 1, 1)
    u = np.stack([1.5 / np.cosh(t) ** 2 + 0j, rng.standard_normal(D) * 0.7 + 0j])
    for scheme in (10, 3):
        d0 = O.akns_degree(scheme)
        tm = np.zeros((2, 4, d0 * D + 1), dtype=np.complex128)
        W = np.zeros(2, dtype=np.int32)
        assert E.emul_fscatter(np.ascontiguousarray(u), None, 2, D, d0, 1, 0, scheme, eps_t, 1, tm, W,
                               use_direct, smem_n) == 0
        for s in range(2):
            tmo, dego, Wo = O.akns_fscatter(u[s], -np.ones(D), eps_t, scheme)
            for e in range(4):
                if np.abs(tmo[e]).sum() > 0:
                    assert rel_err(tm[s, e] * 2.0 ** W[s], tmo[e] * 2.0 ** Wo) < 1e-12
    # explicit r
    q = (rng.standard_normal((1, D)) + 1j * rng.standard_normal((1, D))) * 0.6
    r = (rng.standard_normal((1, D)) + 1j * rng.standard_normal((1, D))) * 0.6
    tm = np.zeros((1, 4, 2 * D + 1), dtype=np.complex128)
    W = np.zeros(1, dtype=np.int32)
    assert E.emul_fscatter(np.ascontiguousarray(q), r.ctypes.data_as(C.c_void_p), 1, D, 2, 2, 0, 10,
                           eps_t, 1, tm, W, use_direct, smem_n) == 0
    tmo, dego, Wo = O.akns_fscatter(q[0], r[0], eps_t, 10)
    for e in range(4):
        assert rel_err(tm[0, e] * 2.0 ** W[0], tmo[e] * 2.0 ** Wo) < 1e-12


@pytest.mark.parametrize("name", sorted(AKNS_TEST_SCHEMES))
@pytest.mark.parametrize("use_direct", [1, 0])
def test_leaf_kernels_all_schemes_reference_golden(E, golden, name, use_direct):
    # the CUDA leaf + tree block programs (host emulation) on the input of
    # test/fnft__akns_fscatter/fnft__akns_fscatter_test_<scheme>.c, explicit r
    q, r, eps_t, z = akns_fscatter_test_input()
    ak = AKNS_TEST_SCHEMES[name]
    d0 = O.akns_degree(ak)
    tm = np.zeros((1, 4, d0 * 8 + 1), dtype=np.complex128)
    W = np.zeros(1, dtype=np.int32)
    r = np.ascontiguousarray(r)
    rc = E.emul_fscatter(np.ascontiguousarray(q), r.ctypes.data, 1, 8, d0, 2, 0, ak, eps_t, 1,
                         tm, W, use_direct, 1024)
    assert rc == 0
    got = eval_tm(tm[0] * 2.0 ** W[0], z)
    assert rel_err(got, golden[f"reftest/akns_fscatter_{name}"]) <= AKNS_TEST_BOUND.get(name, 100) * EPS


@pytest.mark.parametrize("ak", [6, 8, 9, 12, 14, 18])
@pytest.mark.parametrize("D", [3, 37])
def test_chain_leaves_nse_and_kdv_modes_vs_oracle(E, ak, D):
    rng = np.random.default_rng(ak * 100 + D)
    t = np.linspace(-5, 5, D)
    eps_t = 10.0 / (D - 1)
    q = np.stack([1.3 / np.cosh(t) * np.exp(0.7j * t),
                  (rng.standard_normal(D) + 1j * rng.standard_normal(D)) * 0.5])
    d0 = O.akns_degree(ak)
    for rmode, kappa in ((0, 1), (0, -1), (1, 0)):
        tm = np.zeros((2, 4, d0 * D + 1), dtype=np.complex128)
        W = np.zeros(2, dtype=np.int32)
        assert E.emul_fscatter(np.ascontiguousarray(q), None, 2, D, d0, rmode, kappa, ak, eps_t, 1, tm, W,
                               1, 1024) == 0
        for s in range(2):
            r = -kappa * np.conj(q[s]) if rmode == 0 else -np.ones(D)
            tmo, dego, Wo = O.akns_fscatter(q[s], r, eps_t, ak)
            for e in range(4):
                assert rel_err(tm[s, e] * 2.0 ** W[s], tmo[e] * 2.0 ** Wo) < 1e-11

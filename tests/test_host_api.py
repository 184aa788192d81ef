"""CPU-only checks of the drop-in boundary: the shared library loads, exports every
symbol include/fnft_b200.h declares, keeps the reference's struct layouts / defaults /
argument checks, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, ensure_lib

pytestmark = []


@pytest.fixture(scope="module")
def F():
    ensure_lib()
    import fnft_b200
    fnft_b200.lib()
    return fnft_b200


def test_library_exports_every_declared_symbol(F):
    L = F.lib()
    header = open(os.path.join(ROOT, "include", "fnft_b200.h")).read()
    declared = set(re.findall(r"\b(fnft_[a-z0-9_]*|fnft__[a-z0-9_]*)\s*\(", header))
    declared -= {"fnft_printf_ptr_t"}
    assert declared, "no declarations parsed"
    assert set(F.EXPORTED_SYMBOLS) >= declared
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"symbols declared in include/fnft_b200.h but not exported: {missing}"


def test_struct_layouts_match_reference():
    import fnft_b200 as F
    # LP64 layouts verified against the reference with offsetof
    assert C.sizeof(F.NsevOpts) == 48
    assert [getattr(F.NsevOpts, f).offset for f, _ in F.NsevOpts._fields_] == [0, 4, 8, 16, 24, 28, 32, 36, 40]
    assert C.sizeof(F.KdvvOpts) == 4
    assert C.sizeof(F.NsepOpts) == 96
    assert [getattr(F.NsepOpts, f).offset for f, _ in F.NsepOpts._fields_] == [0, 4, 8, 40, 48, 52, 56, 72, 80, 88]


def test_default_options_match_reference(F):
    o = F.nsev_default_opts()            # src/fnft_nsev.c:26-36
    assert (o.bound_state_filtering, o.bound_state_localization, o.niter, o.Dsub) == (2, 2, 10, 0)
    assert (o.discspec_type, o.contspec_type, o.normalization_flag) == (0, 0, 1)
    assert (o.discretization, o.richardson_extrapolation_flag) == (11, 0)
    k = F.kdvv_default_opts()            # src/fnft_kdvv.c:34-36 (2SPLIT8B = 17)
    assert k.discretization == 17
    n = F.nsep_default_opts()            # src/fnft_nsep.c:27-41
    assert (n.localization, n.filtering, n.max_evals, n.discretization) == (2, 2, 20, 4)
    assert (n.normalization_flag, n.points_per_spine, n.Dsub, n.tol) == (1, 2, 0, -1.0)
    assert list(n.floquet_range) == [-1.0, 1.0]
    assert list(n.bounding_box) == [-np.inf, np.inf, -np.inf, np.inf]
    assert F.lib().fnft_nsev_max_K(100, None) == 200  # degree 2 * D


def test_numel_helpers(F):
    L = F.lib()
    # 4*(deg+1)*nextpow2(n), src/private/fnft__poly_fmult.c:40-43
    assert L.fnft__poly_fmult2x2_numel(2, 5) == 4 * 3 * 8
    assert L.fnft__nse_fscatter_numel(100, F.NSE_2SPLIT4B) == 4 * 3 * 128
    assert L.fnft__nse_fscatter_numel(100, F.NSE_BO) == 0          # slow scheme: no polynomial
    assert L.fnft__kdv_fscatter_numel(64, F.KDV_4SPLIT4B) == 4 * 3 * 64
    assert L.fnft__akns_fscatter_numel(64, F.AKNS_2SPLIT2A) == 4 * 2 * 64


def test_version(F):
    L = F.lib()
    a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
    s = C.create_string_buffer(9)
    assert L.fnft_version(C.byref(a), C.byref(b), C.byref(c), s) == 0
    assert (a.value, b.value, c.value, s.value) == (0, 4, 1, b"")


def _call_nsev(F, D=16, q=True, T=(-1.0, 1.0), M=8, XI=(-2.0, 2.0), kappa=1, cs=True, opts=None):
    L = F.lib()
    qa = np.ones(max(D, 1), dtype=np.complex128)
    Ta = np.array(T, dtype=np.float64)
    Xa = np.array(XI, dtype=np.float64)
    out = np.zeros(3 * M, dtype=np.complex128)
    return L.fnft_nsev(D, qa.ctypes.data if q else None, Ta.ctypes.data, M,
                       out.ctypes.data if cs else None, Xa.ctypes.data, None, None, None, kappa,
                       C.addressof(opts) if opts is not None else None)


def test_argument_checks_mirror_reference(F):
    F.lib().fnft_errwarn_setprintf(None)
    EC_INVALID = 2
    # src/fnft_nsev.c:163-174
    assert _call_nsev(F, D=1) == EC_INVALID
    assert _call_nsev(F, q=False) == EC_INVALID
    assert _call_nsev(F, T=(1.0, -1.0)) == EC_INVALID
    assert _call_nsev(F, XI=(2.0, -2.0)) == EC_INVALID
    assert _call_nsev(F, kappa=0) == EC_INVALID
    o = F.nsev_default_opts()
    o.discretization = 99
    assert _call_nsev(F, opts=o) == EC_INVALID
    # src/fnft_kdvv.c:74-92
    L = F.lib()
    u = np.ones(16, dtype=np.complex128)
    Ta = np.array([-1.0, 1.0])
    Xa = np.array([-2.0, 2.0])
    out = np.zeros(8, dtype=np.complex128)
    K = C.c_size_t(1)
    assert L.fnft_kdvv(16, u.ctypes.data, Ta.ctypes.data, 8, None, Xa.ctypes.data, None, None, None, None) == EC_INVALID
    assert L.fnft_kdvv(16, u.ctypes.data, Ta.ctypes.data, 8, out.ctypes.data, Xa.ctypes.data,
                       C.addressof(K), None, None, None) == 6  # NOT_YET_IMPLEMENTED, like the reference


def test_not_yet_implemented_paths_are_loud(F):
    F.lib().fnft_errwarn_setprintf(None)
    o = F.nsev_default_opts()
    o.discretization = 26  # ES4 (like every slow discretization) only with Newton localization: invalid argument
    assert _call_nsev(F, opts=o) == 2
    o = F.nsev_default_opts()
    o.discretization = 1  # BO runs on the GPU, but only with Newton localization (src/fnft_nsev.c:209-219)
    assert _call_nsev(F, opts=o) == 2


def test_no_cpu_fallback_without_gpu(F):
    """On a box without a usable GPU every transform must fail (FNFT_EC_OTHER), never
    silently compute on the CPU."""
    if F.device_count() > 0:
        pytest.skip("a GPU is present")
    msgs = []
    CB = C.CFUNCTYPE(C.c_int32, C.c_char_p)

    F.lib().fnft_errwarn_setprintf(None)
    assert _call_nsev(F) == 5
    u = np.ones(16, dtype=np.complex128)
    ret, tm, deg, W = F.nse_fscatter(u, 0.1, 1, F.NSE_2SPLIT4B)
    assert ret == 5
    ret, out = F.poly_chirpz(np.ones(4), 0.9, np.exp(0.1j), 4)
    assert ret == 5


def test_filter_and_merge_host_logic(F):
    """misc_filter / misc_merge restatements used by the discrete-spectrum driver."""
    from oracle import fnft_oracle as O
    L = F.lib()
    rng = np.random.default_rng(5)
    vals = (rng.uniform(-2, 2, 40) + 1j * rng.uniform(-1, 3, 40)).astype(np.complex128)
    vals[7] = vals[3] + 1e-9
    vals[21] = vals[3] - 1e-10j
    vals[30] = complex(np.nan, 1.0)
    box = np.array([-1.5, 1.5, 0.0, 2.5])
    want = O.misc_merge(O.misc_filter(list(vals), box), np.sqrt(np.finfo(float).eps))
    buf = vals.copy()
    n = C.c_size_t(len(buf))
    L.fnftb__filter_box(C.byref(n), buf.ctypes.data_as(C.c_void_p), box.ctypes.data_as(C.c_void_p))
    L.fnftb__merge.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    L.fnftb__merge(C.byref(n), buf.ctypes.data_as(C.c_void_p), float(np.sqrt(np.finfo(float).eps)))
    assert n.value == len(want)
    assert np.array_equal(buf[:n.value], np.array(want))


def test_logpolar_resolves_unit_circle_defect(F):
    """ln|z| must resolve |z|-1 of a rounded unit-modulus number (needed because the
    chirp-z kernels scale it by n^2/2 ~ 1e9)."""
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    L = F.lib()
    L.fnftb__logpolar.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    for th in [4.8e-6, 0.3, 1.2345, 3.0, -2.2]:
        z = complex(np.exp(1j * th))
        lr, li = C.c_double(), C.c_double()
        L.fnftb__logpolar(z.real, z.imag, C.byref(lr), C.byref(li))
        x, y = Decimal(z.real), Decimal(z.imag)
        exact = float((x * x + y * y).ln() / 2)
        assert abs(lr.value - exact) <= 1e-30 + 1e-12 * abs(exact)
        assert abs(li.value - th) < 1e-15


def test_c_caller_compiles_and_links_against_the_drop_in_headers(F, tmp_path):
    """A C program using only the reference's public names builds against include/ and
    links against libfnft_b200.so (the drop-in claim at the source level)."""
    import subprocess
    exe = str(tmp_path / "nsev_batch_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "nsev_batch_example.c"),
                           "-L" + os.path.join(ROOT, "fnft_b200", "lib"), "-lfnft_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "fnft_b200", "lib"), "-lm", "-o", exe])
    assert os.path.exists(exe)


# ------------------------------------------------------------------ inverse transform: host logic without a GPU
def _inv():
    import inverse_bindings as IB
    return IB


def test_inverse_opts_layout_and_defaults(F):
    # include/fnft_nsev_inverse.h:151-158 (32 bytes on LP64), defaults of src/fnft_nsev_inverse.c:26-33
    IB = _inv()
    assert C.sizeof(IB.InverseOpts) == 32
    assert [getattr(IB.InverseOpts, f).offset for f, _ in IB.InverseOpts._fields_] == [0, 4, 8, 12, 16, 24]
    o = IB.default_opts(F.lib())
    assert (o.discretization, o.contspec_type, o.contspec_inversion_method, o.discspec_type) == (4, 0, 0, 0)
    assert (o.max_iter, o.oversampling_factor) == (100, 8)


def test_inverse_xi_grid_matches_the_reference_formula(F):
    # fnft_nsev_inverse_XI, src/fnft_nsev_inverse.c:40-66: xi of z = exp(2 pi i (M/2+1)/M) and of z = -1
    IB = _inv()
    for D, M, T in ((256, 512, (-2.0, 2.0)), (1024, 1024, (-7.0, 9.0))):
        for disc in (4, 0):
            ret, XI = IB.inverse_XI(F.lib(), D, T, M, disc)
            assert ret == 0
            eps_t = (T[1] - T[0]) / (D - 1)
            z0 = np.exp(2j * np.pi * (M // 2 + 1) / M)
            want = np.array([(np.log(z0) / (2j * eps_t)).real, (np.log(-1 + 0j) / (2j * eps_t)).real])
            assert np.allclose(XI, want, rtol=1e-15, atol=0)
            assert XI[0] < 0 < XI[1]
    assert IB.inverse_XI(F.lib(), 1, (-1.0, 1.0), 8, 4)[0] != 0          # D < 2
    assert IB.inverse_XI(F.lib(), 8, (1.0, -1.0), 8, 4)[0] != 0          # T


def test_inverse_argument_checks_need_no_gpu(F):
    # the checks of src/fnft_nsev_inverse.c:134-172 run before any device work: the same codes with or without a GPU
    IB = _inv()
    F.lib().fnft_errwarn_setprintf(None)
    bs = np.array([0.3 + 1.0j, -0.2 + 0.5j])
    nc = np.array([1.0 + 0j, 2.0 - 1j])
    cs = np.ones(64, dtype=np.complex128)
    XI = np.array([-1.0, 1.0])
    o = IB.default_opts(F.lib())
    E_INVALID, E_SANITY = 2, 7
    assert IB.nsev_inverse(F.lib(), cs[:63], XI, None, None, 32, (-1, 1), 1, o)[0] == E_INVALID     # odd M
    assert IB.nsev_inverse(F.lib(), cs[:16], XI, None, None, 32, (-1, 1), 1, o)[0] == E_INVALID     # M < D
    assert IB.nsev_inverse(F.lib(), cs, XI, None, None, 48, (-1, 1), 1, o)[0] == E_INVALID          # D not 2^k
    assert IB.nsev_inverse(F.lib(), None, None, bs, nc, 32, (1, -1), 1, o)[0] == E_INVALID          # T
    assert IB.nsev_inverse(F.lib(), None, None, bs, nc, 32, (-1, 1), 2, o)[0] == E_INVALID          # kappa
    assert IB.nsev_inverse(F.lib(), None, None, bs, nc, 32, (-1, 1), -1, o)[0] == E_SANITY          # solitons, defocusing
    assert IB.nsev_inverse(F.lib(), None, None, np.conj(bs), nc, 32, (-1, 1), 1, o)[0] == E_SANITY  # lower half plane
    assert IB.nsev_inverse(F.lib(), None, None, None, None, 32, (-1, 1), 1, o)[0] == E_SANITY       # nothing given
    assert IB.nsev_inverse(F.lib(), cs, None, None, None, 32, (-1, 1), 1, o)[0] == E_INVALID        # XI missing
    o.discretization = 11
    assert IB.nsev_inverse(F.lib(), None, None, bs, nc, 32, (-1, 1), 1, o)[0] == E_INVALID          # 2SPLIT4B
    # private symbols: argument checks of fnft__nse_finvscatter (:249-265) and fnft__poly_specfact (:31-38)
    tm = np.zeros(4 * 9, dtype=np.complex128)
    assert IB.nse_finvscatter(F.lib(), tm, -0.1, 1, 4)[0] == E_INVALID      # eps_t
    assert IB.nse_finvscatter(F.lib(), tm, 0.1, 0, 4)[0] == E_INVALID       # kappa
    assert IB.nse_finvscatter(F.lib(), tm, 0.1, 1, 11)[0] != 0              # D = deg / 2 = 4 but not invertible scheme
    assert IB.poly_specfact(F.lib(), np.ones(4, dtype=np.complex128), 0, 1)[0] == E_INVALID   # oversampling 0


def test_nvtx_ranges_are_harmless_without_a_tool():
    # FNFT_B200_NVTX=1 wraps every launch and every public *_batch call in an NVTX range (header-only NVTX3: without an
    # attached tool the calls return at once).  A fresh process, because the variable is read when the library is loaded.
    import subprocess
    import sys
    code = ("import ctypes, os; L = ctypes.CDLL(os.path.join(%r, 'fnft_b200', 'lib', 'libfnft_b200.so'));"
            "L.fnftb_range_push(b'test'); L.fnftb_range_pop();"
            "L.fnft_nsev_batch.restype = ctypes.c_int32;"
            "rc = L.fnft_nsev_batch(ctypes.c_size_t(0), ctypes.c_size_t(0), None, None, ctypes.c_size_t(0), None, None, None,"
            " ctypes.c_size_t(0), None, None, ctypes.c_int32(1), None, None); print('rc', rc)" % ROOT)
    env = dict(os.environ, FNFT_B200_NVTX="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "rc" in out.stdout and "rc 0" not in out.stdout     # invalid arguments are refused, through the ranges


def test_nsev_batch_wrapper_checks_the_caller_owned_spectrum_array(F):
    # contspec_out (fnft_b200.nsev_batch): a caller-owned, pre-touched array for the continuous spectrum; its shape and
    # layout are checked before anything reaches the library
    q = np.zeros((2, 8), dtype=np.complex128)
    for bad in (np.zeros((2, 3), dtype=np.complex128), np.zeros((2, 4), dtype=np.complex64),
                np.zeros((4, 2), dtype=np.complex128).T):
        with pytest.raises(ValueError):
            F.nsev_batch(q, (-1.0, 1.0), 4, (-1.0, 1.0), 1, None, contspec_out=bad)

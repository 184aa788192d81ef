"""ORACLE TOOLING -- ctypes binding to oracle/_ref/libfnft_ref.so.

That shared object is the UNMODIFIED reference (FNFT 0.4.1) compiled by
``oracle/Makefile`` from the sources under /root/reference.  It is test
infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module, and only as
the checker or the timed CPU baseline -- never as part of the product path.

Prototypes follow include/fnft_nsev.h:371-376, include/fnft_kdvv.h:104-109,
include/fnft_nsep.h:263-267 and the private headers under include/private/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libfnft_ref.so")

c_cplx_p = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
c_real_p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class Cplx(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class NsevOpts(C.Structure):
    # include/fnft_nsev.h:198-208 (48 bytes on LP64)
    _fields_ = [
        ("bound_state_filtering", C.c_int),
        ("bound_state_localization", C.c_int),
        ("niter", C.c_size_t),
        ("Dsub", C.c_size_t),
        ("discspec_type", C.c_int),
        ("contspec_type", C.c_int),
        ("normalization_flag", C.c_int32),
        ("discretization", C.c_int),
        ("richardson_extrapolation_flag", C.c_size_t),
    ]


class KdvvOpts(C.Structure):
    # include/fnft_kdvv.h:46-48
    _fields_ = [("discretization", C.c_int)]


class NsepOpts(C.Structure):
    # include/fnft_nsep.h:140-151 (96 bytes on LP64)
    _fields_ = [
        ("localization", C.c_int),
        ("filtering", C.c_int),
        ("bounding_box", C.c_double * 4),
        ("max_evals", C.c_size_t),
        ("discretization", C.c_int),
        ("normalization_flag", C.c_int32),
        ("floquet_range", C.c_double * 2),
        ("points_per_spine", C.c_size_t),
        ("Dsub", C.c_size_t),
        ("tol", C.c_double),
    ]


def available():
    return os.path.exists(REF_SO)


_lib = None


def lib():
    """Load the reference library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError(
            "oracle/_ref/libfnft_ref.so missing: run `make -C oracle ref` in the "
            "build container (needs /root/reference)")
    L = C.CDLL(REF_SO)
    L.fnft_nsev_default_opts.restype = NsevOpts
    L.fnft_kdvv_default_opts.restype = KdvvOpts
    L.fnft_nsep_default_opts.restype = NsepOpts
    L.fnft_nsev.restype = C.c_int32
    L.fnft_nsev.argtypes = [C.c_size_t, C.c_void_p, c_real_p, C.c_size_t, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int32, C.c_void_p]
    L.fnft_kdvv.restype = C.c_int32
    L.fnft_kdvv.argtypes = [C.c_size_t, C.c_void_p, c_real_p, C.c_size_t, C.c_void_p,
                            c_real_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fnft_nsep.restype = C.c_int32
    L.fnft_nsep.argtypes = [C.c_size_t, C.c_void_p, c_real_p, C.c_double, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int32, C.c_void_p]
    L.fnft_nsev_inverse.restype = C.c_int32
    L.fnft_errwarn_setprintf.argtypes = [C.c_void_p]
    L.fnft__nse_fscatter_numel.restype = C.c_size_t
    L.fnft__nse_fscatter_numel.argtypes = [C.c_size_t, C.c_int]
    L.fnft__nse_fscatter.restype = C.c_int32
    L.fnft__nse_fscatter.argtypes = [C.c_size_t, c_cplx_p, C.c_double, C.c_int32,
                                     c_cplx_p, C.c_void_p, C.c_void_p, C.c_int]
    L.fnft__akns_fscatter_numel.restype = C.c_size_t
    L.fnft__akns_fscatter_numel.argtypes = [C.c_size_t, C.c_int]
    L.fnft__akns_fscatter.restype = C.c_int32
    L.fnft__akns_fscatter.argtypes = [C.c_size_t, c_cplx_p, c_cplx_p, C.c_double,
                                      c_cplx_p, C.c_void_p, C.c_void_p, C.c_int]
    L.fnft__kdv_fscatter_numel.restype = C.c_size_t
    L.fnft__kdv_fscatter_numel.argtypes = [C.c_size_t, C.c_int]
    L.fnft__kdv_fscatter.restype = C.c_int32
    L.fnft__kdv_fscatter.argtypes = [C.c_size_t, c_cplx_p, C.c_double, c_cplx_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
    L.fnft__poly_fmult2x2_numel.restype = C.c_size_t
    L.fnft__poly_fmult2x2_numel.argtypes = [C.c_size_t, C.c_size_t]
    L.fnft__poly_fmult2x2.restype = C.c_int32
    L.fnft__poly_fmult2x2.argtypes = [C.c_void_p, C.c_size_t, c_cplx_p, c_cplx_p, C.c_void_p]
    L.fnft__poly_chirpz.restype = C.c_int32
    L.fnft__poly_chirpz.argtypes = [C.c_size_t, c_cplx_p, Cplx, Cplx, C.c_size_t, c_cplx_p]
    L.fnft__nse_scatter_bound_states.restype = C.c_int32
    L.fnft__nse_scatter_bound_states.argtypes = [
        C.c_size_t, c_cplx_p, c_cplx_p, c_real_p, C.c_size_t, c_cplx_p,
        c_cplx_p, c_cplx_p, c_cplx_p, C.c_int, C.c_size_t]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------
# thin numpy-level wrappers
# ---------------------------------------------------------------------------
def nsev_default_opts():
    return lib().fnft_nsev_default_opts()


def nsev(q, T, M=0, XI=None, kappa=+1, opts=None, K=0, bound_states=None,
         want_contspec=True, want_normconsts=True):
    """fnft_nsev (src/fnft_nsev.c:133).  Returns (ret, contspec, K, bs, nc)."""
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.complex128)
    D = q.shape[0]
    T = np.ascontiguousarray(T, dtype=np.float64)
    if opts is None:
        opts = L.fnft_nsev_default_opts()
    cs = None
    if want_contspec and M > 0:
        n = {0: 1, 1: 2, 2: 3}[opts.contspec_type]  # rho, ab, both
        cs = np.zeros(n * M, dtype=np.complex128)
    XIa = None if XI is None else np.ascontiguousarray(XI, dtype=np.float64)
    Kc = C.c_size_t(K)
    bs = nc = None
    if K > 0:
        bs = np.zeros(K, dtype=np.complex128)
        if bound_states is not None:
            bs[:len(bound_states)] = bound_states
        if want_normconsts:
            nc = np.zeros(2 * K, dtype=np.complex128)
    ret = L.fnft_nsev(D, _p(q), T, M, _p(cs), _p(XIa),
                      C.byref(Kc) if K > 0 else None, _p(bs), _p(nc),
                      kappa, C.byref(opts))
    Kout = Kc.value
    return ret, cs, Kout, (None if bs is None else bs[:Kout]), nc


def kdvv(u, T, M, XI, opts=None):
    """fnft_kdvv (src/fnft_kdvv.c:59).  Returns (ret, contspec)."""
    L = lib()
    u = np.ascontiguousarray(u, dtype=np.complex128)
    if opts is None:
        opts = L.fnft_kdvv_default_opts()
    cs = np.zeros(M, dtype=np.complex128)
    ret = L.fnft_kdvv(u.shape[0], _p(u), np.ascontiguousarray(T, dtype=np.float64), M,
                      _p(cs), np.ascontiguousarray(XI, dtype=np.float64),
                      None, None, None, C.byref(opts))
    return ret, cs


def nsep(q, T, kappa=+1, opts=None, K=None, M=None, phase_shift=0.0):
    """fnft_nsep (src/fnft_nsep.c:82).  Returns (ret, main_spec, aux_spec)."""
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.complex128)
    D = q.shape[0]
    if opts is None:
        opts = L.fnft_nsep_default_opts()
    K = 64 * D if K is None else K
    M = 64 * D if M is None else M
    main = np.zeros(K, dtype=np.complex128)
    aux = np.zeros(M, dtype=np.complex128)
    Kc, Mc = C.c_size_t(K), C.c_size_t(M)
    ret = L.fnft_nsep(D, _p(q), np.ascontiguousarray(T, dtype=np.float64), phase_shift,
                      C.byref(Kc), _p(main), C.byref(Mc), _p(aux), None, kappa,
                      C.byref(opts))
    return ret, main[:Kc.value], aux[:Mc.value]


def nse_fscatter(q, eps_t, kappa, discretization, normalize=True):
    """fnft__nse_fscatter (src/private/fnft__nse_fscatter.c:44).
    Returns (ret, tm[4, deg+1], deg, W)."""
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.complex128)
    D = q.shape[0]
    numel = L.fnft__nse_fscatter_numel(D, discretization)
    res = np.zeros(numel, dtype=np.complex128)
    deg = C.c_size_t(0)
    W = C.c_int32(0)
    ret = L.fnft__nse_fscatter(D, q, eps_t, kappa, res, C.byref(deg),
                               C.byref(W) if normalize else None, discretization)
    d = deg.value
    return ret, res[:4 * (d + 1)].reshape(4, d + 1).copy(), d, W.value


def akns_fscatter(q, r, eps_t, discretization, normalize=True):
    """fnft__akns_fscatter (src/private/fnft__akns_fscatter.c:64)."""
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.complex128)
    r = np.ascontiguousarray(r, dtype=np.complex128)
    D = q.shape[0]
    numel = L.fnft__akns_fscatter_numel(D, discretization)
    res = np.zeros(numel, dtype=np.complex128)
    deg = C.c_size_t(0)
    W = C.c_int32(0)
    ret = L.fnft__akns_fscatter(D, q, r, eps_t, res, C.byref(deg),
                                C.byref(W) if normalize else None, discretization)
    d = deg.value
    return ret, res[:4 * (d + 1)].reshape(4, d + 1).copy(), d, W.value


def kdv_fscatter(u, eps_t, discretization, normalize=True):
    """fnft__kdv_fscatter (src/private/fnft__kdv_fscatter.c:45)."""
    L = lib()
    u = np.ascontiguousarray(u, dtype=np.complex128)
    D = u.shape[0]
    numel = L.fnft__kdv_fscatter_numel(D, discretization)
    res = np.zeros(numel, dtype=np.complex128)
    deg = C.c_size_t(0)
    W = C.c_int32(0)
    ret = L.fnft__kdv_fscatter(D, u, eps_t, res, C.byref(deg),
                               C.byref(W) if normalize else None, discretization)
    d = deg.value
    return ret, res[:4 * (d + 1)].reshape(4, d + 1).copy(), d, W.value


def poly_fmult2x2(deg, p, normalize=True):
    """fnft__poly_fmult2x2 (src/private/fnft__poly_fmult.c:381).
    p: [4, n, deg+1] (entry-major, highest power first).
    Returns (ret, result[4, deg_out+1], deg_out, W)."""
    L = lib()
    p = np.ascontiguousarray(p, dtype=np.complex128)
    n = p.shape[1]
    numel = L.fnft__poly_fmult2x2_numel(deg, n)
    buf = np.zeros(numel, dtype=np.complex128)
    buf[:p.size] = p.reshape(-1)
    res = np.zeros(numel, dtype=np.complex128)
    d = C.c_size_t(deg)
    W = C.c_int32(0)
    ret = L.fnft__poly_fmult2x2(C.byref(d), n, buf, res,
                                C.byref(W) if normalize else None)
    do = d.value
    return ret, res[:4 * (do + 1)].reshape(4, do + 1).copy(), do, W.value


def poly_chirpz(p, A, W, M):
    """fnft__poly_chirpz (src/private/fnft__poly_chirpz.c:33)."""
    L = lib()
    p = np.ascontiguousarray(p, dtype=np.complex128)
    out = np.zeros(M, dtype=np.complex128)
    A = complex(A)
    W = complex(W)
    ret = L.fnft__poly_chirpz(p.shape[0] - 1, p, Cplx(A.real, A.imag),
                              Cplx(W.real, W.imag), M, out)
    return ret, out


def nse_scatter_bound_states(q, r, T, lam, discretization, skip_b=False):
    """fnft__nse_scatter_bound_states
    (src/private/fnft__nse_scatter_bound_states.c:29). Returns (ret, a, aprime, b)."""
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.complex128)
    r = np.ascontiguousarray(r, dtype=np.complex128)
    lam = np.ascontiguousarray(lam, dtype=np.complex128)
    K = lam.shape[0]
    a = np.zeros(K, dtype=np.complex128)
    ap = np.zeros(K, dtype=np.complex128)
    b = np.zeros(K, dtype=np.complex128)
    ret = L.fnft__nse_scatter_bound_states(q.shape[0], q, r,
                                           np.ascontiguousarray(T, dtype=np.float64),
                                           K, lam, a, ap, b, discretization,
                                           1 if skip_b else 0)
    return ret, a, ap, b

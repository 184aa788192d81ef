"""ctypes bindings of the inverse-transform entry points, usable with the drop-in library (fnft_b200.lib()) and with
the unmodified reference (oracle.ref_lib.lib()): both export the same C signatures
(include/fnft_nsev_inverse.h:168-263, include/private/fnft__nse_finvscatter.h:61-63, fnft__poly_specfact.h:62-66)."""
import ctypes as C

import numpy as np


class InverseOpts(C.Structure):
    _fields_ = [("discretization", C.c_int), ("contspec_type", C.c_int), ("contspec_inversion_method", C.c_int),
                ("discspec_type", C.c_int), ("max_iter", C.c_size_t), ("oversampling_factor", C.c_size_t)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_opts(L):
    L.fnft_nsev_inverse_default_opts.restype = InverseOpts
    return L.fnft_nsev_inverse_default_opts()


def inverse_XI(L, D, T, M, disc):
    XI = np.zeros(2)
    Ta = np.ascontiguousarray(T, dtype=np.float64)
    L.fnft_nsev_inverse_XI.argtypes = None
    ret = L.fnft_nsev_inverse_XI(C.c_size_t(D), _p(Ta), C.c_size_t(M), _p(XI), C.c_int(disc))
    return ret, XI


def nsev_inverse(L, contspec, XI, bound_states, normconsts, D, T, kappa, opts):
    """returns (ret, q, contspec as modified by the call)"""
    cs = None if contspec is None else np.ascontiguousarray(contspec, dtype=np.complex128).copy()
    M = 0 if cs is None else cs.shape[0]
    bs = None if bound_states is None else np.ascontiguousarray(bound_states, dtype=np.complex128)
    nc = None if normconsts is None else np.ascontiguousarray(normconsts, dtype=np.complex128)
    K = 0 if bs is None else bs.shape[0]
    q = np.zeros(D, dtype=np.complex128)
    Ta = np.ascontiguousarray(T, dtype=np.float64)
    XIa = None if XI is None else np.ascontiguousarray(XI, dtype=np.float64)
    L.fnft_nsev_inverse.argtypes = None
    L.fnft_nsev_inverse.restype = C.c_int32
    ret = L.fnft_nsev_inverse(C.c_size_t(M), _p(cs), _p(XIa), C.c_size_t(K), _p(bs), _p(nc), C.c_size_t(D), _p(q),
                              _p(Ta), C.c_int32(kappa), C.byref(opts))
    return ret, q, cs


def nsev_inverse_batch(L, contspec, XI, bound_states, normconsts, D, T, kappa, opts):
    """drop-in library only.  contspec [B][M] or None, bound_states / normconsts [B][K] or None"""
    cs = None if contspec is None else np.ascontiguousarray(contspec, dtype=np.complex128).copy()
    bs = None if bound_states is None else np.ascontiguousarray(bound_states, dtype=np.complex128)
    nc = None if normconsts is None else np.ascontiguousarray(normconsts, dtype=np.complex128)
    B = cs.shape[0] if cs is not None else bs.shape[0]
    M = 0 if cs is None else cs.shape[1]
    K = 0 if bs is None else bs.shape[1]
    q = np.zeros((B, D), dtype=np.complex128)
    rcs = np.zeros(B, dtype=np.int32)
    Ta = np.ascontiguousarray(T, dtype=np.float64)
    XIa = None if XI is None else np.ascontiguousarray(XI, dtype=np.float64)
    L.fnft_nsev_inverse_batch.argtypes = None
    L.fnft_nsev_inverse_batch.restype = C.c_int32
    ret = L.fnft_nsev_inverse_batch(C.c_size_t(B), C.c_size_t(M), _p(cs), _p(XIa), C.c_size_t(K), _p(bs), _p(nc),
                                    C.c_size_t(D), _p(q), _p(Ta), C.c_int32(kappa), C.byref(opts), _p(rcs))
    return ret, q, rcs


def nse_finvscatter(L, tm, eps_t, kappa, disc):
    """tm: [4][deg+1] (reference layout); returns (ret, q[deg])"""
    tm = np.ascontiguousarray(tm, dtype=np.complex128).copy()
    deg = tm.size // 4 - 1
    q = np.zeros(deg, dtype=np.complex128)
    L.fnft__nse_finvscatter.argtypes = None
    L.fnft__nse_finvscatter.restype = C.c_int32
    ret = L.fnft__nse_finvscatter(C.c_size_t(deg), _p(tm), _p(q), C.c_double(eps_t), C.c_int32(kappa), C.c_int(disc))
    return ret, q


def poly_specfact(L, poly, oversampling, kappa):
    poly = np.ascontiguousarray(poly, dtype=np.complex128)
    deg = poly.shape[0] - 1
    res = np.zeros(deg + 1, dtype=np.complex128)
    L.fnft__poly_specfact.argtypes = None
    L.fnft__poly_specfact.restype = C.c_int32
    ret = L.fnft__poly_specfact(C.c_size_t(deg), _p(poly), _p(res), C.c_size_t(oversampling), C.c_int32(kappa))
    return ret, res

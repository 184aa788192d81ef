"""fnft_b200 -- Python (ctypes) face of libfnft_b200.so.

The product is the C-ABI shared library ``fnft_b200/lib/libfnft_b200.so`` (C host
code + hand-written sm_100a CUDA kernels, see ``include/fnft_b200.h``).  This module
only loads it and offers numpy-level wrappers whose names, argument order and
meaning mirror the reference's C interface (``fnft_nsev``, ``fnft_kdvv``,
``fnft__nse_fscatter`` ...), so that the parity tests read like the reference's own
tests.  There is no Python or CPU implementation of the hot path in here: if the
library is missing, or no CUDA device is usable, every call fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FNFT_B200_LIB") or os.path.join(_HERE, "lib", "libfnft_b200.so")

# enum values (include/fnft_b200.h)
NSE_2SPLIT2_MODAL, NSE_BO, NSE_2SPLIT1A, NSE_2SPLIT1B, NSE_2SPLIT2A, NSE_2SPLIT2B, NSE_2SPLIT2S = range(7)
NSE_2SPLIT4B = 11
NSE_4SPLIT4B = 21
NSE_CF4_2 = 22
KDV_2SPLIT1A, KDV_2SPLIT1B, KDV_2SPLIT2A, KDV_2SPLIT2B, KDV_2SPLIT2S = range(5)
KDV_2SPLIT4B = 9
KDV_4SPLIT4B = 19
AKNS_2SPLIT2_MODAL, AKNS_2SPLIT1A, AKNS_2SPLIT1B, AKNS_2SPLIT2A, AKNS_2SPLIT2B, AKNS_2SPLIT2S = range(6)
AKNS_2SPLIT4B = 10
AKNS_4SPLIT4B = 21
BSFILT_NONE, BSFILT_BASIC, BSFILT_FULL = range(3)
BSLOC_FAST_EIGENVALUE, BSLOC_NEWTON, BSLOC_SUBSAMPLE_AND_REFINE = range(3)
DSTYPE_NORMING_CONSTANTS, DSTYPE_RESIDUES, DSTYPE_BOTH = range(3)
CSTYPE_REFLECTION_COEFFICIENT, CSTYPE_AB, CSTYPE_BOTH = range(3)
NSEP_LOC_SUBSAMPLE_AND_REFINE, NSEP_LOC_GRIDSEARCH, NSEP_LOC_MIXED = range(3)
NSEP_FILT_NONE, NSEP_FILT_MANUAL, NSEP_FILT_AUTO = range(3)


class Cplx(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class NsevOpts(C.Structure):
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
    _fields_ = [("discretization", C.c_int)]


class NsepOpts(C.Structure):
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


# every symbol include/fnft_b200.h declares
EXPORTED_SYMBOLS = [
    "fnft_errwarn_setprintf", "fnft_errwarn_getprintf", "fnft_version",
    "fnft_nsev_default_opts", "fnft_nsev_max_K", "fnft_nsev",
    "fnft_kdvv_default_opts", "fnft_kdvv",
    "fnft_nsep_default_opts", "fnft_nsep",
    "fnft__poly_fmult2x2_numel", "fnft__poly_fmult2x2", "fnft__poly_chirpz", "fnft__poly_roots_fasteigen",
    "fnft__akns_fscatter_numel", "fnft__akns_fscatter",
    "fnft__nse_fscatter_numel", "fnft__nse_fscatter",
    "fnft__kdv_fscatter_numel", "fnft__kdv_fscatter",
    "fnft__nse_scatter_bound_states",
    "fnft_nsev_batch", "fnft_kdvv_batch", "fnft_nsep_batch",
    "fnft_b200_device_count", "fnft_b200_set_device", "fnft_b200_set_device_pointers",
    "fnft_b200_synchronize", "fnft_b200_set_workspace_limit", "fnft_b200_stream",
    "fnft_b200_launch_count", "fnft_b200_release",
    "fnft_b200_profile_enable", "fnft_b200_profile_report",
    "fnft_b200_set_devices", "fnft_b200_get_devices", "fnft_b200_probe_fp64_tflops",
    "fnft__errmsg_aux", "fnft__warn_aux",
    "fnft__poly_fmult_numel", "fnft__poly_fmult", "fnft__poly_eval", "fnft__poly_evalderiv",
    "fnft__misc_print_buf", "fnft__misc_rel_err", "fnft__misc_hausdorff_dist", "fnft__misc_sech",
    "fnft__misc_l2norm2", "fnft__misc_filter", "fnft__misc_filter_inv", "fnft__misc_filter_nonreal",
    "fnft__misc_merge", "fnft__misc_downsample", "fnft__misc_CSINC", "fnft__misc_nextpowerof2",
    "fnft_nsev_inverse_default_opts", "fnft_nsev_inverse_XI", "fnft_nsev_inverse", "fnft_nsev_inverse_batch",
    "fnft__nse_finvscatter", "fnft__poly_specfact",
]

_lib = None


def lib():
    """Load libfnft_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
            "fnft_b200 has no Python/CPU implementation of the hot path")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, dbl = C.c_void_p, C.c_size_t, C.c_int32, C.c_double
    L.fnft_nsev_default_opts.restype = NsevOpts
    L.fnft_kdvv_default_opts.restype = KdvvOpts
    L.fnft_nsev_max_K.restype = sz
    L.fnft_nsev_max_K.argtypes = [sz, vp]
    L.fnft_nsev.restype = i32
    L.fnft_nsev.argtypes = [sz, vp, vp, sz, vp, vp, vp, vp, vp, i32, vp]
    L.fnft_nsev_batch.restype = i32
    L.fnft_nsev_batch.argtypes = [sz, sz, vp, vp, sz, vp, vp, vp, sz, vp, vp, i32, vp, vp]
    L.fnft_kdvv.restype = i32
    L.fnft_kdvv.argtypes = [sz, vp, vp, sz, vp, vp, vp, vp, vp, vp]
    L.fnft_kdvv_batch.restype = i32
    L.fnft_kdvv_batch.argtypes = [sz, sz, vp, vp, sz, vp, vp, vp, vp]
    if hasattr(L, "fnft_nsep"):
        L.fnft_nsep_default_opts.restype = NsepOpts
        L.fnft_nsep.restype = i32
        L.fnft_nsep.argtypes = [sz, vp, vp, dbl, vp, vp, vp, vp, vp, i32, vp]
        L.fnft_nsep_batch.restype = i32
        L.fnft_nsep_batch.argtypes = [sz, sz, vp, vp, dbl, vp, sz, vp, vp, sz, vp, i32, vp, vp]
    L.fnft__poly_fmult2x2_numel.restype = sz
    L.fnft__poly_fmult2x2_numel.argtypes = [sz, sz]
    L.fnft__poly_fmult2x2.restype = i32
    L.fnft__poly_fmult2x2.argtypes = [vp, sz, vp, vp, vp]
    L.fnft__poly_fmult_numel.restype = sz
    L.fnft__poly_fmult_numel.argtypes = [sz, sz]
    L.fnft__poly_fmult.restype = i32
    L.fnft__poly_fmult.argtypes = [vp, sz, vp, vp]
    L.fnft__poly_eval.restype = i32
    L.fnft__poly_eval.argtypes = [sz, vp, sz, vp]
    L.fnft__poly_evalderiv.restype = i32
    L.fnft__poly_evalderiv.argtypes = [sz, vp, sz, vp, vp]
    L.fnft__misc_rel_err.restype = dbl
    L.fnft__misc_rel_err.argtypes = [i32, vp, vp]
    L.fnft__misc_hausdorff_dist.restype = dbl
    L.fnft__misc_hausdorff_dist.argtypes = [sz, vp, sz, vp]
    L.fnft__misc_l2norm2.restype = dbl
    L.fnft__misc_l2norm2.argtypes = [sz, vp, dbl, dbl]
    L.fnft__misc_filter.restype = i32
    L.fnft__misc_filter.argtypes = [vp, vp, vp, vp]
    L.fnft__misc_filter_inv.restype = i32
    L.fnft__misc_filter_inv.argtypes = [vp, vp, vp, vp]
    L.fnft__misc_filter_nonreal.restype = i32
    L.fnft__misc_filter_nonreal.argtypes = [vp, vp, dbl]
    L.fnft__misc_merge.restype = i32
    L.fnft__misc_merge.argtypes = [vp, vp, dbl]
    L.fnft__misc_nextpowerof2.restype = sz
    L.fnft__misc_nextpowerof2.argtypes = [sz]
    L.fnft__poly_roots_fasteigen.restype = i32
    L.fnft__poly_roots_fasteigen.argtypes = [sz, vp, vp]
    L.fnft__poly_chirpz.restype = i32
    L.fnft__poly_chirpz.argtypes = [sz, vp, Cplx, Cplx, sz, vp]
    L.fnft__akns_fscatter_numel.restype = sz
    L.fnft__akns_fscatter_numel.argtypes = [sz, C.c_int]
    L.fnft__akns_fscatter.restype = i32
    L.fnft__akns_fscatter.argtypes = [sz, vp, vp, dbl, vp, vp, vp, C.c_int]
    L.fnft__nse_fscatter_numel.restype = sz
    L.fnft__nse_fscatter_numel.argtypes = [sz, C.c_int]
    L.fnft__nse_fscatter.restype = i32
    L.fnft__nse_fscatter.argtypes = [sz, vp, dbl, i32, vp, vp, vp, C.c_int]
    L.fnft__kdv_fscatter_numel.restype = sz
    L.fnft__kdv_fscatter_numel.argtypes = [sz, C.c_int]
    L.fnft__kdv_fscatter.restype = i32
    L.fnft__kdv_fscatter.argtypes = [sz, vp, dbl, vp, vp, vp, C.c_int]
    L.fnft__nse_scatter_bound_states.restype = i32
    L.fnft__nse_scatter_bound_states.argtypes = [sz, vp, vp, vp, sz, vp, vp, vp, vp, C.c_int, sz]
    L.fnft_errwarn_setprintf.argtypes = [vp]
    L.fnft_b200_device_count.restype = i32
    L.fnft_b200_set_device.restype = i32
    L.fnft_b200_set_device.argtypes = [i32]
    L.fnft_b200_set_device_pointers.argtypes = [i32]
    L.fnft_b200_synchronize.restype = i32
    L.fnft_b200_set_workspace_limit.argtypes = [sz]
    L.fnft_b200_stream.restype = vp
    L.fnft_b200_launch_count.restype = C.c_ulonglong
    L.fnft_b200_profile_enable.argtypes = [i32]
    L.fnft_b200_profile_report.restype = C.c_char_p
    L.fnft_b200_set_devices.restype = i32
    L.fnft_b200_set_devices.argtypes = [i32, vp]
    L.fnft_b200_get_devices.restype = i32
    L.fnft_b200_get_devices.argtypes = [vp, i32]
    L.fnft_b200_probe_fp64_tflops.restype = dbl
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    return int(lib().fnft_b200_device_count())


def set_device(dev):
    return int(lib().fnft_b200_set_device(int(dev)))


def set_devices(devs):
    """Several GPUs behind one *_batch call (host buffers): shard i of the batch runs on devs[i]."""
    a = np.ascontiguousarray(list(devs), dtype=np.int32)
    return int(lib().fnft_b200_set_devices(len(a), _p(a) if len(a) else None))


def launch_count():
    return int(lib().fnft_b200_launch_count())


def quiet(flag=True):
    """Silence (or restore) the calling thread's FNFT error/warning printing."""
    L = lib()
    if flag:
        L.fnft_errwarn_setprintf(None)
    else:
        raise NotImplementedError("restoring the default printf needs the original pointer")


def nsev_default_opts():
    return lib().fnft_nsev_default_opts()


def kdvv_default_opts():
    return lib().fnft_kdvv_default_opts()


def nsep_default_opts():
    return lib().fnft_nsep_default_opts()


_CS_LEN = {CSTYPE_REFLECTION_COEFFICIENT: 1, CSTYPE_AB: 2, CSTYPE_BOTH: 3}


def nsev(q, T, M=0, XI=None, kappa=+1, opts=None, K=0, bound_states=None,
         want_contspec=True, want_normconsts=True):
    """fnft_nsev (include/fnft_nsev.h:371-376).  Same convention as
    oracle.ref_lib.nsev: returns (ret, contspec, K, bound_states, normconsts)."""
    L = lib()
    q = _c128(q)
    D = q.shape[0]
    T = _f64(T)
    if opts is None:
        opts = L.fnft_nsev_default_opts()
    cs = None
    if want_contspec and M > 0:
        cs = np.zeros(_CS_LEN[opts.contspec_type] * M, dtype=np.complex128)
    XIa = None if XI is None else _f64(XI)
    Kc = C.c_size_t(K)
    bs = nc = None
    if K > 0:
        bs = np.zeros(K, dtype=np.complex128)
        if bound_states is not None:
            bs[:len(bound_states)] = bound_states
        if want_normconsts:
            nc = np.zeros(2 * K, dtype=np.complex128)
    ret = L.fnft_nsev(D, _p(q), _p(T), M, _p(cs), _p(XIa),
                      C.addressof(Kc) if K > 0 else None, _p(bs), _p(nc), kappa,
                      C.addressof(opts))
    Kout = Kc.value
    return ret, cs, Kout, (None if bs is None else bs[:Kout]), nc


def nsev_batch(q, T, M=0, XI=None, kappa=+1, opts=None, K=None, Kmax=0, bound_states=None,
               want_normconsts=True, contspec_out=None):
    """fnft_nsev_batch (include/fnft_b200.h).  q: [B, D].  Returns
    (ret, contspec[B, len] or None, K[B] or None, bound_states[B, Kmax] or None,
     normconsts[B, nlen] or None, ret_codes[B]).  contspec_out: a caller-owned (e.g. pinned, already touched)
    complex128 array [B, len] that receives the continuous spectrum instead of a fresh np.zeros array."""
    L = lib()
    q = _c128(q)
    B, D = q.shape
    T = _f64(T)
    if opts is None:
        opts = L.fnft_nsev_default_opts()
    cs = None
    if M > 0 and XI is not None:
        if contspec_out is not None:
            cs = contspec_out
            if cs.dtype != np.complex128 or not cs.flags.c_contiguous or cs.shape != (B, _CS_LEN[opts.contspec_type] * M):
                raise ValueError("contspec_out must be a C-contiguous complex128 array [B, len]")
        else:
            cs = np.zeros((B, _CS_LEN[opts.contspec_type] * M), dtype=np.complex128)
    XIa = None if XI is None else _f64(XI)
    Ka = bs = nc = None
    if Kmax > 0:
        Ka = np.ascontiguousarray(K, dtype=np.uint64).copy()
        bs = np.zeros((B, Kmax), dtype=np.complex128)
        bs[:, :bound_states.shape[1]] = bound_states
        if want_normconsts:
            nlen = 2 * Kmax if opts.discspec_type == DSTYPE_BOTH else Kmax
            nc = np.zeros((B, nlen), dtype=np.complex128)
    rcs = np.zeros(B, dtype=np.int32)
    ret = L.fnft_nsev_batch(B, D, _p(q), _p(T), M, _p(cs), _p(XIa), _p(Ka), Kmax, _p(bs), _p(nc),
                            kappa, C.addressof(opts), _p(rcs))
    return ret, cs, Ka, bs, nc, rcs


def kdvv(u, T, M, XI, opts=None):
    """fnft_kdvv (include/fnft_kdvv.h:104-109).  Returns (ret, contspec)."""
    L = lib()
    u = _c128(u)
    if opts is None:
        opts = L.fnft_kdvv_default_opts()
    cs = np.zeros(M, dtype=np.complex128)
    ret = L.fnft_kdvv(u.shape[0], _p(u), _p(_f64(T)), M, _p(cs), _p(_f64(XI)), None, None, None,
                      C.addressof(opts))
    return ret, cs


def kdvv_batch(u, T, M, XI, opts=None):
    """fnft_kdvv_batch.  u: [B, D].  Returns (ret, contspec[B, M], ret_codes[B])."""
    L = lib()
    u = _c128(u)
    B, D = u.shape
    if opts is None:
        opts = L.fnft_kdvv_default_opts()
    cs = np.zeros((B, M), dtype=np.complex128)
    rcs = np.zeros(B, dtype=np.int32)
    ret = L.fnft_kdvv_batch(B, D, _p(u), _p(_f64(T)), M, _p(cs), _p(_f64(XI)), C.addressof(opts),
                            _p(rcs))
    return ret, cs, rcs


def nsep(q, T, kappa=+1, opts=None, K=None, M=None, phase_shift=0.0):
    """fnft_nsep (include/fnft_nsep.h:263-267).  Returns (ret, main_spec, aux_spec)."""
    L = lib()
    q = _c128(q)
    D = q.shape[0]
    if opts is None:
        opts = L.fnft_nsep_default_opts()
    K = 64 * D if K is None else K
    M = 64 * D if M is None else M
    main = np.zeros(K, dtype=np.complex128)
    aux = np.zeros(M, dtype=np.complex128)
    Kc, Mc = C.c_size_t(K), C.c_size_t(M)
    ret = L.fnft_nsep(D, _p(q), _p(_f64(T)), phase_shift, C.addressof(Kc), _p(main),
                      C.addressof(Mc), _p(aux), None, kappa, C.addressof(opts))
    return ret, main[:Kc.value], aux[:Mc.value]


def nsep_buffers(B, Kmax, Mmax):
    """Output arrays of fnft_nsep_batch, allocated AND touched (a fresh np.zeros array is mapped lazily: the first
    write into every 2 MiB region costs a page fault + zero fill, ~55 ms for the 2 x 268 MB of BASELINE config 5,
    which a timed call would otherwise pay).  Pass them as `out` to nsep_batch and reuse them."""
    bufs = (np.zeros(B, dtype=np.uint64), np.empty((B, Kmax), dtype=np.complex128), np.zeros(B, dtype=np.uint64),
            np.empty((B, Mmax), dtype=np.complex128), np.zeros(B, dtype=np.int32))
    bufs[1].fill(0)
    bufs[3].fill(0)
    return bufs


def nsep_batch(q, T, Kmax, Mmax, kappa=+1, opts=None, phase_shift=0.0, out=None):
    """fnft_nsep_batch.  Returns (ret, K[B], main[B,Kmax], Mcount[B], aux[B,Mmax], rcs); out = nsep_buffers(...)
    reuses the caller's arrays (only the first K[b] / Mcount[b] entries of a row are written, like the reference)."""
    L = lib()
    q = _c128(q)
    B, D = q.shape
    if opts is None:
        opts = L.fnft_nsep_default_opts()
    if out is None:
        Ka = np.zeros(B, dtype=np.uint64)
        Ma = np.zeros(B, dtype=np.uint64)
        main = np.zeros((B, Kmax), dtype=np.complex128)
        aux = np.zeros((B, Mmax), dtype=np.complex128)
        rcs = np.zeros(B, dtype=np.int32)
    else:
        Ka, main, Ma, aux, rcs = out
        assert main.shape == (B, Kmax) and aux.shape == (B, Mmax) and Ka.shape == (B,) and Ma.shape == (B,)
    ret = L.fnft_nsep_batch(B, D, _p(q), _p(_f64(T)), phase_shift, _p(Ka), Kmax, _p(main), _p(Ma),
                            Mmax, _p(aux), kappa, C.addressof(opts), _p(rcs))
    return ret, Ka, main, Ma, aux, rcs


def _fscatter_result(ret, res, deg, W):
    d = deg.value
    return ret, res[:4 * (d + 1)].reshape(4, d + 1).copy(), d, W.value


def nse_fscatter(q, eps_t, kappa, discretization, normalize=True):
    """fnft__nse_fscatter.  Returns (ret, tm[4, deg+1], deg, W)."""
    L = lib()
    q = _c128(q)
    D = q.shape[0]
    res = np.zeros(L.fnft__nse_fscatter_numel(D, discretization), dtype=np.complex128)
    deg, W = C.c_size_t(0), C.c_int32(0)
    ret = L.fnft__nse_fscatter(D, _p(q), eps_t, kappa, _p(res), C.addressof(deg),
                               C.addressof(W) if normalize else None, discretization)
    return _fscatter_result(ret, res, deg, W)


def akns_fscatter(q, r, eps_t, discretization, normalize=True):
    """fnft__akns_fscatter.  Returns (ret, tm[4, deg+1], deg, W)."""
    L = lib()
    q, r = _c128(q), _c128(r)
    D = q.shape[0]
    res = np.zeros(L.fnft__akns_fscatter_numel(D, discretization), dtype=np.complex128)
    deg, W = C.c_size_t(0), C.c_int32(0)
    ret = L.fnft__akns_fscatter(D, _p(q), _p(r), eps_t, _p(res), C.addressof(deg),
                                C.addressof(W) if normalize else None, discretization)
    return _fscatter_result(ret, res, deg, W)


def kdv_fscatter(u, eps_t, discretization, normalize=True):
    """fnft__kdv_fscatter.  Returns (ret, tm[4, deg+1], deg, W)."""
    L = lib()
    u = _c128(u)
    D = u.shape[0]
    res = np.zeros(L.fnft__kdv_fscatter_numel(D, discretization), dtype=np.complex128)
    deg, W = C.c_size_t(0), C.c_int32(0)
    ret = L.fnft__kdv_fscatter(D, _p(u), eps_t, _p(res), C.addressof(deg),
                               C.addressof(W) if normalize else None, discretization)
    return _fscatter_result(ret, res, deg, W)


def poly_fmult2x2(deg, p, normalize=True):
    """fnft__poly_fmult2x2.  p: [4, n, deg+1].  Returns (ret, result[4, deg_out+1], deg_out, W)."""
    L = lib()
    p = _c128(p)
    n = p.shape[1]
    numel = L.fnft__poly_fmult2x2_numel(deg, n)
    buf = np.zeros(numel, dtype=np.complex128)
    buf[:p.size] = p.reshape(-1)
    res = np.zeros(numel, dtype=np.complex128)
    d, W = C.c_size_t(deg), C.c_int32(0)
    ret = L.fnft__poly_fmult2x2(C.addressof(d), n, _p(buf), _p(res),
                                C.addressof(W) if normalize else None)
    do = d.value
    return ret, res[:4 * (do + 1)].reshape(4, do + 1).copy(), do, W.value


def poly_chirpz(p, A, W, M):
    """fnft__poly_chirpz.  Returns (ret, result[M])."""
    L = lib()
    p = _c128(p)
    out = np.zeros(M, dtype=np.complex128)
    A, W = complex(A), complex(W)
    ret = L.fnft__poly_chirpz(p.shape[0] - 1, _p(p), Cplx(A.real, A.imag), Cplx(W.real, W.imag), M,
                              _p(out))
    return ret, out


def poly_roots_fasteigen(p):
    """fnft__poly_roots_fasteigen.  p: deg+1 coefficients, highest power first.
    Returns (ret, roots[deg])."""
    L = lib()
    p = _c128(p)
    roots = np.zeros(p.shape[0] - 1, dtype=np.complex128)
    ret = L.fnft__poly_roots_fasteigen(p.shape[0] - 1, _p(p), _p(roots))
    return ret, roots


def nse_scatter_bound_states(q, r, T, lam, discretization, skip_b=False):
    """fnft__nse_scatter_bound_states.  Returns (ret, a, aprime, b)."""
    L = lib()
    q = _c128(q)
    r = None if r is None else _c128(r)
    lam = _c128(lam)
    K = lam.shape[0]
    a = np.zeros(K, dtype=np.complex128)
    ap = np.zeros(K, dtype=np.complex128)
    b = np.zeros(K, dtype=np.complex128)
    ret = L.fnft__nse_scatter_bound_states(q.shape[0], _p(q), _p(r), _p(_f64(T)), K, _p(lam), _p(a),
                                           _p(ap), _p(b), discretization, 1 if skip_b else 0)
    return ret, a, ap, b

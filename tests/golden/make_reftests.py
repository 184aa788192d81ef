"""Generates tests/golden/reftests.npz + reftests.json -- run in the BUILD CONTAINER only.

The reference's end-to-end tests (test/fnft_nsev/*.c, test/fnft_kdvv/*.c, test/fnft_nsep/*.c) all have the form
    opts = defaults; opts.<field> = ...; <x>_testcases_test_fnft(tc, D, error_bounds, &opts); ...
Each test file is compiled here UNCHANGED from /root/reference together with a stub of
<x>_testcases_test_fnft that only records its arguments (test case, D, the six error bounds, the
option struct), which yields the exact list of calls the reference's test suite makes -- without
copying or parsing any reference source.  The test-case data (signal, exact spectra) come from the
reference's own generator functions fnft__nsev_testcases / fnft__kdvv_testcases in oracle/_ref.
tests/test_reference_suite.py replays every recorded call against libfnft_b200.so on the GPU and
applies the reference's own error bounds.

    python tests/golden/make_reftests.py
"""
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_lib as R  # noqa: E402

REF = "/root/reference"
REFLIB_DIR = os.path.join(ROOT, "oracle", "_ref")

STUB = r'''
#include <stdio.h>
#include "fnft_nsev.h"
#include "fnft_kdvv.h"
#include "fnft_nsep.h"
FNFT_INT fnft__nsep_testcases_test_fnft(int tc, FNFT_UINT D, FNFT_REAL eb[3], fnft_nsep_opts_t *o)
{
    fnft_nsep_opts_t d = fnft_nsep_default_opts();
    if (o == NULL)
        o = &d;
    printf("{\"fn\": \"nsep\", \"tc\": %d, \"D\": %zu, \"eb\": [%.17g, %.17g, %.17g], \"localization\": %d, "
           "\"filtering\": %d, \"bounding_box\": [%.17g, %.17g, %.17g, %.17g], \"max_evals\": %zu, "
           "\"discretization\": %d, \"normalization_flag\": %d, \"floquet_range\": [%.17g, %.17g], "
           "\"points_per_spine\": %zu, \"Dsub\": %zu, \"tol\": %.17g}\n", tc, D, eb[0], eb[1], eb[2],
           (int)o->localization, (int)o->filtering, o->bounding_box[0], o->bounding_box[1], o->bounding_box[2],
           o->bounding_box[3], o->max_evals, (int)o->discretization, (int)o->normalization_flag,
           o->floquet_range[0], o->floquet_range[1], o->points_per_spine, o->Dsub, o->tol);
    return 0;
}
FNFT_INT fnft__nsev_testcases_test_fnft(int tc, FNFT_UINT D, const FNFT_REAL eb[6], fnft_nsev_opts_t *o)
{
    printf("{\"fn\": \"nsev\", \"tc\": %d, \"D\": %zu, \"eb\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], "
           "\"bsfilt\": %d, \"bsloc\": %d, \"niter\": %zu, \"Dsub\": %zu, \"normalization_flag\": %d, "
           "\"discretization\": %d, \"richardson\": %d}\n", tc, D, eb[0], eb[1], eb[2], eb[3], eb[4], eb[5],
           (int)o->bound_state_filtering, (int)o->bound_state_localization, o->niter, o->Dsub,
           (int)o->normalization_flag, (int)o->discretization, (int)o->richardson_extrapolation_flag);
    return 0;
}
FNFT_INT fnft__kdvv_testcases_test_fnft(int tc, FNFT_UINT D, const FNFT_REAL eb[6], fnft_kdvv_opts_t *o)
{
    printf("{\"fn\": \"kdvv\", \"tc\": %d, \"D\": %zu, \"eb\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], "
           "\"discretization\": %d}\n", tc, D, eb[0], eb[1], eb[2], eb[3], eb[4], eb[5], (int)o->discretization);
    return 0;
}
'''


def record_calls():
    calls = []
    with tempfile.TemporaryDirectory() as tmp:
        stub = os.path.join(tmp, "stub.c")
        open(stub, "w").write(STUB)
        inc = ["-I" + REFLIB_DIR, "-I" + os.path.join(REF, "include"), "-I" + os.path.join(REF, "include", "private"),
               "-I" + os.path.join(REF, "include", "3rd_party", "kiss_fft")]
        for sub in ("fnft_nsev", "fnft_kdvv", "fnft_nsep"):
            for src in sorted(glob.glob(os.path.join(REF, "test", sub, "*.c"))):
                exe = os.path.join(tmp, "t.out")
                subprocess.check_call(["gcc", "-std=gnu99", "-w", "-O0"] + inc + [src, stub, "-L" + REFLIB_DIR,
                                      "-lfnft_ref", "-lm", "-Wl,-rpath," + REFLIB_DIR, "-o", exe])
                out = subprocess.run([exe], capture_output=True, text=True, timeout=600).stdout
                for line in out.splitlines():
                    line = line.strip().replace("inf", "Infinity").replace("nan", "NaN")
                    if line.startswith("{"):
                        c = json.loads(line)
                        c["file"] = os.path.basename(src)
                        calls.append(c)
    return calls


def nsev_case(tc, D):
    L = R.lib()
    q, cs, ab, bs, nc, res = (C.c_void_p() for _ in range(6))
    T, XI = (C.c_double * 2)(), (C.c_double * 2)()
    M, K, kappa = C.c_size_t(), C.c_size_t(), C.c_int32()
    f = L.fnft__nsev_testcases
    f.restype = C.c_int32
    f.argtypes = None
    rc = f(C.c_int(tc), C.c_size_t(D), C.byref(q), T, C.byref(M), C.byref(cs), C.byref(ab), XI, C.byref(K),
           C.byref(bs), C.byref(nc), C.byref(res), C.byref(kappa))
    assert rc == 0

    def arr(p, n):
        if not p.value or n == 0:
            return np.zeros(0, dtype=np.complex128)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128).copy()
    Mv, Kv = M.value, K.value
    return dict(q=arr(q, D), T=np.array(T[:]), XI=np.array(XI[:]), kappa=np.array(kappa.value),
                contspec=arr(cs, Mv), ab=arr(ab, 2 * Mv), bound_states=arr(bs, Kv), normconsts=arr(nc, Kv),
                residues=arr(res, Kv))


def kdvv_case(tc, D):
    L = R.lib()
    q, cs, ab, bs, nc, res = (C.c_void_p() for _ in range(6))
    T, XI = (C.c_double * 2)(), (C.c_double * 2)()
    M, K = C.c_size_t(), C.c_size_t()
    f = L.fnft__kdvv_testcases
    f.restype = C.c_int32
    f.argtypes = None
    rc = f(C.c_int(tc), C.c_size_t(D), C.byref(q), T, C.byref(M), C.byref(cs), C.byref(ab), XI, C.byref(K),
           C.byref(bs), C.byref(nc), C.byref(res))
    assert rc == 0
    Mv = M.value
    a = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128).copy()
    return dict(q=a(q, D), T=np.array(T[:]), XI=np.array(XI[:]), contspec=a(cs, Mv))


def nsep_case(tc, D):
    L = R.lib()
    q, ms, au, sh = (C.c_void_p() for _ in range(4))
    T, rb = (C.c_double * 2)(), (C.c_double * 4)()
    ps = C.c_double()
    K, M, kappa = C.c_size_t(), C.c_size_t(), C.c_int32()
    f = L.fnft__nsep_testcases
    f.restype = C.c_int32
    f.argtypes = None
    rc = f(C.c_int(tc), C.c_size_t(D), C.byref(q), T, C.byref(ps), C.byref(K), C.byref(ms), C.byref(M), C.byref(au),
           C.byref(sh), C.byref(kappa), rb)
    assert rc == 0
    a = lambda p, n: (np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128).copy()
                      if p.value and n else np.zeros(0, dtype=np.complex128))
    return dict(q=a(q, D), T=np.array(T[:]), phase_shift=np.array(ps.value), kappa=np.array(kappa.value),
                mainspec=a(ms, K.value), auxspec=a(au, M.value), remove_box=np.array(rb[:]))


def main():
    calls = record_calls()
    G = {}
    for c in calls:
        key = "%s/%d/%d" % (c["fn"], c["tc"], c["D"])
        if key + "/q" in G:
            continue
        data = {"nsev": nsev_case, "kdvv": kdvv_case, "nsep": nsep_case}[c["fn"]](c["tc"], c["D"])
        for k, v in data.items():
            G[key + "/" + k] = v
    np.savez_compressed(os.path.join(HERE, "reftests.npz"), **G)
    json.dump(calls, open(os.path.join(HERE, "reftests.json"), "w"), indent=0)
    print("recorded %d calls from %d files; %d arrays, %d bytes" % (
        len(calls), len({c["file"] for c in calls}), len(G), os.path.getsize(os.path.join(HERE, "reftests.npz"))))


if __name__ == "__main__":
    main()

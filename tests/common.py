"""Shared inputs of the parity tests: the formulas of the reference's own unit tests
(test/fnft__poly/*.c, test/fnft__akns_fscatter/*.c) and seeded synthetic signals."""
import numpy as np


def rel_err(num, exact):
    """misc_rel_err (src/private/fnft__misc.c:41-51): sum|num-exact| / sum|exact|."""
    num, exact = np.asarray(num), np.asarray(exact)
    return np.abs(num - exact).sum() / np.abs(exact).sum()


def fmult2x2_test_input(n, deg=1):
    """Input of fnft__poly_fmult2x2_test_n_is_(no_)power_of_2.c:(96-99): p[4, n, deg+1]."""
    i = np.arange((deg + 1) * n, dtype=np.float64)
    p = np.empty((4, (deg + 1) * n), dtype=np.complex128)
    for e, off in enumerate((0.0, 0.1, 0.2, 0.3)):
        p[e] = np.sqrt(i + 1.0) * (np.cos(i + off) + 1j * np.sin(-2.0 * i + off))
    return p.reshape(4, n, deg + 1)


def akns_fscatter_test_input():
    """q, r, eps_t, z of test/fnft__akns_fscatter/fnft__akns_fscatter_test_*.c:37-38,101-102."""
    i = np.arange(1, 9, dtype=np.float64)
    q = (0.41 * np.cos(i) + 0.59j * np.sin(0.28 * i)) * 50
    r = (0.33 * np.sin(i) + 0.85j * np.cos(0.43 * i)) * 25
    z = np.array([1.0, np.exp(1j * np.pi / 4), np.exp(1j * 9 * np.pi / 14),
                  np.exp(1j * 4 * np.pi / 3), np.exp(-1j * np.pi / 5)])
    return q, r, 0.13, z


AKNS_TEST_SCHEMES = {"2split4B": 10, "2split2A": 3, "2split1A": 1, "2split1B": 2, "2split2B": 4,
                     "2split2S": 5, "2split2_modal": 0, "2split3A": 6, "2split3B": 7, "2split3S": 8,
                     "2split4A": 9, "2split5A": 11, "2split5B": 12, "2split6A": 13, "2split6B": 14,
                     "2split7A": 15, "2split7B": 16, "2split8A": 17, "2split8B": 18}
# error bounds of the reference tests in units of eps (err_bnd in the files named above)
AKNS_TEST_BOUND = {"2split6A": 291, "2split6B": 250, "2split7A": 250, "2split7B": 250,
                   "2split8A": 250, "2split8B": 250}


def eval_tm(tm, z):
    """poly_eval of the four entries at the points z (Horner, highest power first)."""
    return np.concatenate([np.polyval(tm[e], z) for e in range(4)])


CHIRPZ_TEST_P = np.array([1 + 2j, -3 - 0.5j, 0.3, -0.4j])
CHIRPZ_TEST_A = 0.95
CHIRPZ_TEST_W = np.exp(0.3j)


def sech_chirp(D, T, amp=2.0, chirp=0.3):
    t = np.linspace(T[0], T[1], D)
    return amp / np.cosh(t) * np.exp(1j * chirp * t * t)


def ld_evaluator(eps_t, step_div):
    """evaluate(p, xi) for oracle.nsev_contspec / oracle.kdvv: Horner's rule in 80-bit long double
    at z = exp(2i xi eps_t / step_div) -- the exact value of the polynomial the double-precision
    chirp-z implementations approximate (the reference's loses accuracy where sum|c| >> |p(z)|)."""
    def ev(p, xi):
        zz = np.exp(2j * xi.astype(np.longdouble) * np.longdouble(eps_t) / step_div)
        r = np.zeros_like(zz)
        for ck in p.astype(np.clongdouble):
            r = r * zz + ck
        return r
    return ev


def parity_contract(ours, ref, tol=1e-9, floor=1e-2):
    """The parity contract of this repository ("Parity contract"), derived
    from SURVEY.md 8(c).  Returns two figures normalised by `tol`; both must be < 1:
      (i)  misc_rel_err(ours, ref) = sum|d| / sum|ref|                  (the reference's
           own comparison metric, src/private/fnft__misc.c:41-51)
      (ii) pointwise  |d| <= tol * (|ref| + floor * max|ref|)
    The additive term of (ii) is the reference's OWN accuracy floor: its chirp-z forms
    the chirp with cpow(W, n^2/2), whose phase error (~1e-16 * n^2/2 * arg W rad) gives
    every output an absolute error of order 1e-12..1e-11 * max|.|.  Where |ref| is small
    the reference is therefore only accurate to ~1e-8 relative (demonstrated against a
    long-double evaluation in tests/test_oracle.py::test_reference_accuracy_floor), and a
    purely relative 1e-9 bound against it is not a meaningful target there."""
    ours, ref = np.asarray(ours), np.asarray(ref)
    mx = np.abs(ref).max()
    e1 = rel_err(ours, ref)
    e2 = (np.abs(ours - ref) / (np.abs(ref) + floor * mx)).max()
    return e1 / tol, e2 / tol


def parity_pointwise(ours, ref, tol=1e-9, split=1e-6):
    """SURVEY.md 8(c) (ii) and (iii) exactly as written: pointwise |d| <= tol * |ref| on the points with
    |ref| >= split * max|ref|, and |d| <= tol * max|ref| on the remaining (tail) points.  Returns the worst
    ratio normalised by tol (must be < 1).  Attainable only where the reference's own absolute error floor
    (cpow-based chirp, ~1e-11 * max) stays below tol * |ref|; the tests that assert it say so."""
    ours, ref = np.asarray(ours), np.asarray(ref)
    a = np.abs(ref)
    mx = a.max()
    d = np.abs(ours - ref)
    big = a >= split * mx
    e = 0.0
    if big.any():
        e = max(e, float((d[big] / a[big]).max()))
    if (~big).any():
        e = max(e, float(d[~big].max() / mx))
    return e / tol


# ---------------------------------------------------------------------------------
# build helpers (the product library is built in-tree by `make`; tests build it on
# demand if it is missing so that the CPU suite is self-contained)
# ---------------------------------------------------------------------------------
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ensure_lib():
    path = os.path.join(ROOT, "fnft_b200", "lib", "libfnft_b200.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-j8"], cwd=ROOT)
    return path


def ensure_emul():
    path = os.path.join(ROOT, "tests", "emul", "libfnftb_emul.so")
    srcs = [os.path.join(ROOT, "tests", "emul", "emul_lib.cpp")]
    cud = os.path.join(ROOT, "fnft_b200", "csrc", "cuda")
    srcs += [os.path.join(cud, f) for f in os.listdir(cud) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "emul"], cwd=ROOT)
    return path

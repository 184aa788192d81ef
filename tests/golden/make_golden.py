"""Generates tests/golden/golden.npz -- run in the BUILD CONTAINER only.

Two sources (both need /root/reference, which does not exist on the GPU box):
  1. the literal golden arrays of the reference's own unit tests, parsed out of
     /root/reference/test/**.c (keys "reftest/...");
  2. outputs of the unmodified reference library (oracle/_ref/libfnft_ref.so, built by
     oracle/Makefile) on seeded inputs (keys "refrun/...").
The committed .npz travels to the GPU box; the tests never read /root/reference.

    python tests/golden/make_golden.py
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_lib as R  # noqa: E402

REF = "/root/reference"


def parse_complex_array(path, name):
    """Extracts `COMPLEX name[..] = { a + b*I, ... };` from a C file."""
    src = open(path).read()
    m = re.search(r"COMPLEX\s+" + re.escape(name) + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", src, re.S)
    if m is None:
        raise KeyError(f"{name} not found in {path}")
    body = m.group(1).replace("\\", " ")
    vals = []
    for item in body.split(","):
        s = item.strip()
        if not s:
            continue
        s = s.replace("I*", "1j*").replace("*I", "*1j")
        s = re.sub(r"(?<![\w.])I(?![\w.])", "1j", s)
        vals.append(complex(eval(s, {"__builtins__": {}}, {})))
    return np.array(vals, dtype=np.complex128)


def main():
    G = {}
    # ---------------------------------------------------------------- 1. reference tests
    t = os.path.join(REF, "test")
    G["reftest/fmult2x2_pow2"] = parse_complex_array(
        os.path.join(t, "fnft__poly/fnft__poly_fmult2x2_test_n_is_power_of_2.c"), "result_exact")
    G["reftest/fmult2x2_nopow2"] = parse_complex_array(
        os.path.join(t, "fnft__poly/fnft__poly_fmult2x2_test_n_is_no_power_of_2.c"), "result_exact")
    G["reftest/chirpz_M3"] = parse_complex_array(
        os.path.join(t, "fnft__poly/fnft__poly_chirpz_test.c"), "result_exactM3")
    G["reftest/chirpz_M6"] = parse_complex_array(
        os.path.join(t, "fnft__poly/fnft__poly_chirpz_test.c"), "result_exactM6")
    for nm in ["2split4B", "2split2A", "2split1A", "2split1B", "2split2B", "2split2S", "2split2_modal",
               "2split3A", "2split3B", "2split3S", "2split4A", "2split5A", "2split5B", "2split6A",
               "2split6B", "2split7A", "2split7B", "2split8A", "2split8B"]:
        G[f"reftest/akns_fscatter_{nm}"] = parse_complex_array(
            os.path.join(t, f"fnft__akns_fscatter/fnft__akns_fscatter_test_{nm}.c"), "result_exact")

    # ---------------------------------------------------------------- 2. reference runs
    rng = np.random.default_rng(20261018)

    def sig(D, kind):
        tt = np.linspace(-6, 6, D)
        if kind == 0:
            return 1.9 / np.cosh(tt) * np.exp(0.3j * tt * tt + 0.2j)
        return (rng.standard_normal(D) + 1j * rng.standard_normal(D)) * 0.4 * np.exp(-tt * tt / 8)

    for disc in (11, 4):
        for D in (2, 3, 37, 64, 100, 256):
            q = sig(D, D % 2)
            eps = 12.0 / max(D - 1, 1)
            ret, tm, deg, W = R.nse_fscatter(q, eps, 1, disc)
            assert ret == 0
            G[f"refrun/fscatter/{disc}/{D}/q"] = q
            G[f"refrun/fscatter/{disc}/{D}/eps"] = np.array(eps)
            G[f"refrun/fscatter/{disc}/{D}/tm"] = tm * 2.0 ** W
    R.lib().fnft_errwarn_setprintf(None)
    for disc in (11, 4, 21):
        for D in ((37, 64, 256, 1000) if disc != 21 else (64, 256)):
            for kappa in (1, -1):
                q = sig(D, 0) if disc == 21 else sig(D, (D // 2) % 2)
                o = R.nsev_default_opts()
                o.discretization = disc
                o.contspec_type = 2
                ret, cs, *_ = R.nsev(q, [-6, 6], 32, [-3.5, 2.75], kappa, o)
                assert ret == 0
                G[f"refrun/nsev/{disc}/{D}/{kappa}/q"] = q
                G[f"refrun/nsev/{disc}/{D}/{kappa}/cs"] = cs
    for disc in (9, 19, 2):
        for D in (64, 100):
            tt = np.linspace(-16, 15, D)
            u = 1.7 / np.cosh(tt) ** 2 + 0.3 * np.exp(-(tt - 1) ** 2)
            o = R.lib().fnft_kdvv_default_opts()
            o.discretization = disc
            ret, cs = R.kdvv(u, [-16, 15], 32, [-3.55, 3.95], o)
            assert ret == 0
            G[f"refrun/kdvv/{disc}/{D}/u"] = u.astype(np.complex128)
            G[f"refrun/kdvv/{disc}/{D}/cs"] = cs
    D = 256
    tt = np.linspace(-12, 12, D)
    q = 2.8 / np.cosh(tt) * np.exp(0.3j * tt)
    g = np.array([0.35j - 0.14, 1.25j - 0.17, 2.2j - 0.1, 0.31j - 0.15])
    for disc in (11, 21):
        o = R.nsev_default_opts()
        o.discretization = disc
        o.bound_state_localization = 1
        o.discspec_type = 2
        ret, cs, K, bs, nc = R.nsev(q, [-12, 12], 0, None, 1, o, K=4, bound_states=g,
                                    want_contspec=False)
        assert ret == 0
        G[f"refrun/bound/{disc}/q"] = q
        G[f"refrun/bound/{disc}/guesses"] = g
        G[f"refrun/bound/{disc}/bs"] = bs
        G[f"refrun/bound/{disc}/nc"] = nc[:2 * K]
    # a, a', b of the BO recurrence at fixed points (kernel-level pin)
    lam = np.array([0.4j + 0.1, 1.3j - 0.2, 2.0j])
    ret, a, ap, b = R.nse_scatter_bound_states(q, -np.conj(q), [-12, 12], lam, 1)
    assert ret == 0
    G["refrun/scatter_bo/q"] = q
    G["refrun/scatter_bo/lam"] = lam
    G["refrun/scatter_bo/a"] = a
    G["refrun/scatter_bo/ap"] = ap
    G["refrun/scatter_bo/b"] = b
    p = rng.standard_normal(41) + 1j * rng.standard_normal(41)
    A, W = 0.97 * np.exp(0.2j), 1.0005 * np.exp(0.03j)
    ret, out = R.poly_chirpz(p, A, W, 25)
    G["refrun/chirpz/p"] = p
    G["refrun/chirpz/AW"] = np.array([A, W])
    G["refrun/chirpz/out"] = out
    pm = rng.standard_normal((4, 5, 4)) + 1j * rng.standard_normal((4, 5, 4))
    ret, res, deg, W = R.poly_fmult2x2(3, pm)
    assert ret == 0
    G["refrun/fmult2x2_deg3_n5/p"] = pm
    G["refrun/fmult2x2_deg3_n5/res"] = res * 2.0 ** W
    # every polynomial splitting scheme end to end: nsev (a, b, rho) and kdvv (R).  The
    # reference's chirp-z loses accuracy where sum|coeff| >> |p(z)| (high-degree schemes,
    # kappa = -1), so next to its output ("cs") the same quantities are recorded with the
    # oracle's polynomials (pinned to the reference's to 1e-14) evaluated by Horner's rule in
    # 80-bit long double ("exact").
    from oracle import fnft_oracle as O

    def ld_eval(eps_t, step_div):
        def ev(p, xi):
            zz = np.exp(2j * xi.astype(np.longdouble) * np.longdouble(eps_t) / step_div)
            r = np.zeros_like(zz)
            for ck in p.astype(np.clongdouble):
                r = r * zz + ck
            return r
        return ev

    for disc in (2, 3, 5, 6, 7, 8, 9, 10, 12, 13, 14, 15, 16, 17, 18, 19, 20):   # nse enum values
        D = 64 if disc == 20 else 50
        tt = np.linspace(-6, 6, D)
        q = 1.6 / np.cosh(tt) * np.exp(0.25j * tt * tt - 0.4j * tt)
        for kappa in (1, -1):
            o = R.nsev_default_opts()
            o.discretization = disc
            o.contspec_type = 2
            ret, cs, *_ = R.nsev(q, [-6, 6], 24, [-2.5, 3.25], kappa, o)
            assert ret == 0, (disc, ret)
            G[f"refrun/schemes_nsev/{disc}/{kappa}/q"] = q
            G[f"refrun/schemes_nsev/{disc}/{kappa}/cs"] = cs
            sch = O._NSE2AKNS[disc]
            G[f"refrun/schemes_nsev/{disc}/{kappa}/exact"] = O.nsev_contspec(
                q, [-6, 6], 24, [-2.5, 3.25], kappa, disc, cstype=2,
                evaluate=ld_eval(12.0 / (D - 1), O.akns_degree(sch) * O.akns_upsampling(sch))).astype(np.complex128)
    for disc in range(0, 19):   # kdv enum values 1A ... 4SPLIT4A (4SPLIT4B = 19 is above)
        D = 64 if disc == 18 else 50
        tt = np.linspace(-16, 15, D)
        u = 1.3 / np.cosh(tt) ** 2 + 0.25 * np.exp(-(tt + 2) ** 2)
        o = R.lib().fnft_kdvv_default_opts()
        o.discretization = disc
        ret, cs = R.kdvv(u, [-16, 15], 24, [-3.55, 3.95], o)
        assert ret == 0, (disc, ret)
        G[f"refrun/schemes_kdvv/{disc}/u"] = u.astype(np.complex128)
        G[f"refrun/schemes_kdvv/{disc}/cs"] = cs
        G[f"refrun/schemes_kdvv/{disc}/exact"] = O.kdvv(
            u, [-16, 15], 24, [-3.55, 3.95], disc,
            evaluate=ld_eval(31.0 / (D - 1), O.akns_degree(O._KDV2AKNS[disc]))).astype(np.complex128)
    # default options (SUBSAMPLE_AND_REFINE) and FAST_EIGENVALUE: the reference with eiscor replaced
    # by the oracle-only companion-matrix shim (oracle/eiscor_shim.c)
    def bs_case(key, q, T, M, XI, disc, bsloc, dstype=2):
        o = R.nsev_default_opts()
        o.discretization = disc
        o.bound_state_localization = bsloc
        o.discspec_type = dstype
        ret, cs, K, bs, nc = R.nsev(q, T, M, XI, 1, o, K=2 * len(q))
        assert ret == 0, (key, ret)
        G[f"refrun/defaults/{key}/q"] = np.asarray(q, dtype=np.complex128)
        G[f"refrun/defaults/{key}/par"] = np.array([T[0], T[1], M, XI[0], XI[1], disc, bsloc, dstype], dtype=np.float64)
        G[f"refrun/defaults/{key}/cs"] = cs
        G[f"refrun/defaults/{key}/bs"] = bs[:K]
        G[f"refrun/defaults/{key}/nc"] = nc[:2 * K if dstype == 2 else K]
    # BASELINE config 1: examples/fnft_nsev_example.c as shipped (rectangular pulse, :34-76)
    bs_case("example", np.full(256, 2.0 + 0j), [-1, 1], 8, [-2, 2], 11, 2, 0)
    for D in (256, 500, 1024):
        tt = np.linspace(-10, 10, D)
        qs = 2.7 / np.cosh(tt) * np.exp(0.4j * tt)
        bs_case(f"sech{D}_sub", qs, [-10, 10], 16, [-2, 2], 11, 2)
        bs_case(f"sech{D}_fast", qs, [-10, 10], 16, [-2, 2], 11, 0)
    tt = np.linspace(-10, 10, 256)
    qs = 2.7 / np.cosh(tt) * np.exp(0.4j * tt)
    bs_case("sech256_4split4b_sub", qs, [-10, 10], 16, [-2, 2], 21, 2)
    bs_case("sech256_2split2a_sub", qs, [-10, 10], 16, [-2, 2], 4, 2)
    bs_case("sech256_2split6b_fast", qs, [-10, 10], 16, [-2, 2], 15, 0)
    # Richardson extrapolation (src/fnft_nsev.c:316-442)
    def re_case(key, q, T, M, XI, disc, bsloc, dstype, cstype, guesses=None):
        o = R.nsev_default_opts()
        o.discretization, o.bound_state_localization, o.discspec_type = disc, bsloc, dstype
        o.contspec_type, o.richardson_extrapolation_flag = cstype, 1
        if guesses is None:
            ret, cs, K, bs, nc = R.nsev(q, T, M, XI, 1, o, K=2 * len(q))
        else:
            ret, cs, K, bs, nc = R.nsev(q, T, M, XI, 1, o, K=len(guesses), bound_states=guesses)
        assert ret == 0, (key, ret)
        G[f"refrun/richardson/{key}/q"] = np.asarray(q, dtype=np.complex128)
        G[f"refrun/richardson/{key}/par"] = np.array([T[0], T[1], M, XI[0], XI[1], disc, bsloc, dstype, cstype], dtype=np.float64)
        G[f"refrun/richardson/{key}/guesses"] = np.zeros(0, dtype=np.complex128) if guesses is None else guesses
        G[f"refrun/richardson/{key}/cs"] = cs
        G[f"refrun/richardson/{key}/bs"] = bs[:K]
        G[f"refrun/richardson/{key}/nc"] = nc[:2 * K if dstype == 2 else K]
    gs = np.array([0.2j - 0.2, 1.2j - 0.21, 2.19j - 0.2])
    for D in (256, 301):
        tt = np.linspace(-10, 10, D)
        qs = 2.7 / np.cosh(tt) * np.exp(0.4j * tt)
        re_case(f"sech{D}_newton_both", qs, [-10, 10], 16, [-2, 2], 11, 1, 2, 2, gs)
        re_case(f"sech{D}_newton_res", qs, [-10, 10], 16, [-2, 2], 11, 1, 1, 0, gs)
        re_case(f"sech{D}_sub_nc", qs, [-10, 10], 16, [-2, 2], 11, 2, 0, 1)
        re_case(f"sech{D}_sub_res", qs, [-10, 10], 16, [-2, 2], 11, 2, 1, 0)
        re_case(f"sech{D}_2a_sub", qs, [-10, 10], 16, [-2, 2], 4, 2, 2, 2)
    tt = np.linspace(-10, 10, 256)
    qs = 2.7 / np.cosh(tt) * np.exp(0.4j * tt)
    re_case("sech256_4split4b", qs, [-10, 10], 16, [-2, 2], 21, 1, 2, 2, gs)
    # fnft_nsep with the default options (MIXED, 2SPLIT2A) and SUBSAMPLE_AND_REFINE: perturbed plane
    # waves (the family of BASELINE config 5), reference with the companion-matrix stand-in for eiscor
    def nsep_case(key, D, A, m, k, e, phi, kappa, loc, disc, phase_shift=0.0):
        tt = 2 * np.pi * np.arange(D) / D
        qq = A * np.exp(1j * m * tt) * (1 + e * np.cos(k * tt + phi))
        if phase_shift != 0.0:
            qq = qq * np.exp(1j * phase_shift * tt / (2 * np.pi))
        o = R.lib().fnft_nsep_default_opts()
        o.localization, o.discretization = loc, disc
        ret, ms, au = R.nsep(qq, [0, 2 * np.pi], kappa, o, phase_shift=phase_shift)
        assert ret == 0, (key, ret)
        G[f"refrun/nsep_defaults/{key}/q"] = qq
        G[f"refrun/nsep_defaults/{key}/par"] = np.array([kappa, loc, disc, phase_shift], dtype=np.float64)
        G[f"refrun/nsep_defaults/{key}/main"] = ms
        G[f"refrun/nsep_defaults/{key}/aux"] = au
    nsep_case("mixed_2a", 256, 1.3, 1, 2, 0.2, 0.3, 1, 2, 4)
    nsep_case("sub_2a", 256, 1.3, 1, 2, 0.2, 0.3, 1, 0, 4)
    nsep_case("mixed_4b", 256, 0.9, 0, 1, 0.25, 1.1, 1, 2, 11)
    nsep_case("sub_4b_defoc", 128, 0.9, 0, 1, 0.25, 1.1, -1, 0, 11)
    nsep_case("mixed_2a_defoc", 128, 1.1, 2, 3, 0.1, 0.0, -1, 2, 4)
    nsep_case("mixed_2a_shift", 256, 1.3, 1, 2, 0.2, 0.3, 1, 2, 4, phase_shift=0.8)
    nsep_case("sub_4split4b", 128, 1.2, 0, 1, 0.2, 0.5, 1, 0, 21)
    nsep_case("mixed_2a_1024", 1024, 1.6, 0, 2, 0.15, 0.4, 1, 2, 4)
    # slow discretizations BO (1) and CF4_2 (22): continuous spectrum by nse_scatter_matrix + Newton
    for disc in (1, 22):
        for D, kappa in ((100, 1), (256, 1), (256, -1)):
            tt = np.linspace(-10, 10, D)
            qs = 2.7 / np.cosh(tt) * np.exp(0.4j * tt)
            o = R.nsev_default_opts()
            o.discretization, o.bound_state_localization, o.discspec_type, o.contspec_type = disc, 1, 2, 2
            gs = np.array([0.2j - 0.2, 1.2j - 0.21, 2.19j - 0.2])
            ret, cs, K, bs, nc = R.nsev(qs, [-10, 10], 20, [-2, 2.5], kappa, o, K=3, bound_states=gs)
            assert ret == 0, (disc, D, kappa, ret)
            key = f"refrun/slow/{disc}/{D}/{kappa}"
            G[key + "/q"] = qs
            G[key + "/guesses"] = gs
            G[key + "/cs"] = cs
            G[key + "/bs"] = bs[:K]
            G[key + "/nc"] = nc[:2 * K]
    pr = rng.standard_normal(60) + 1j * rng.standard_normal(60)
    G["refrun/roots/p"] = pr
    ret = R.lib().fnft__poly_roots_fasteigen
    rts = np.zeros(59, dtype=np.complex128)
    assert ret(59, pr.ctypes.data, rts.ctypes.data) == 0
    G["refrun/roots/roots"] = rts
    G["reftest/roots_fasteigen"] = parse_complex_array(
        os.path.join(t, "fnft__poly/fnft__poly_roots_fasteigen_test.c"), "roots_exact")
    # accuracy floor of the reference: defocusing case, D = 126, evaluated exactly
    # (leaf coefficients in double like every implementation, product tree by direct
    # convolution and Horner evaluation in 80-bit long double)
    from oracle import fnft_oracle as O
    Dq, Mq, kap = 126, 40, -1
    Tq, XIq = (-32.0, 32.0), (-10.0, 10.0)
    tq = np.linspace(Tq[0], Tq[1], Dq)
    qq = 1.7 / np.cosh(tq) * np.exp(-3j * tq + 0.4j * np.sin(tq))
    eps_t = (Tq[1] - Tq[0]) / (Dq - 1)
    leaves = O.akns_leaves(qq, -kap * np.conj(qq), eps_t, 10)
    LD = np.clongdouble

    def conv(a, b):
        o = np.zeros(len(a) + len(b) - 1, dtype=LD)
        for i, ai in enumerate(a):
            o[i:i + len(b)] += ai * b
        return o
    mats = [[leaves[e, k].astype(LD) for e in range(4)] for k in range(Dq)]
    while len(mats) > 1:
        nxt = []
        for i in range(0, len(mats) - 1, 2):
            A, Bm = mats[i], mats[i + 1]
            nxt.append([conv(A[0], Bm[0]) + conv(A[1], Bm[2]), conv(A[0], Bm[1]) + conv(A[1], Bm[3]),
                        conv(A[2], Bm[0]) + conv(A[3], Bm[2]), conv(A[2], Bm[1]) + conv(A[3], Bm[3])])
        if len(mats) % 2:
            nxt.append(mats[-1])
        mats = nxt
    eps_xi = (XIq[1] - XIq[0]) / (Mq - 1)
    xi = XIq[0] + eps_xi * np.arange(Mq)
    z = np.exp(2j * xi.astype(np.longdouble) * np.longdouble(eps_t) / 2)

    def horner(cf, zz):
        r = np.zeros_like(zz)
        for ck in cf:
            r = r * zz + ck
        return r
    rho_exact = (horner(mats[0][2], z) * np.exp(1j * xi * (-2 * (Tq[1] + eps_t * 0.5)))
                 / horner(mats[0][0], z)).astype(np.complex128)
    o = R.nsev_default_opts()
    ret, cs, *_ = R.nsev(qq, Tq, Mq, XIq, kap, o)
    assert ret == 0
    G["floor/q"] = qq
    G["floor/rho_exact"] = rho_exact
    G["floor/rho_reference"] = cs
    out = os.path.join(HERE, "golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3, 4 and 5 (bench.py measures config 2):

  3  fnft_nsev bound states + norming constants (Newton), D = 4096, 8-soliton signals, B = 1024
  4  fnft_kdvv reflection coefficient, 4SPLIT4B, D = M = 8192, B = 2048
  5  fnft_nsep main + auxiliary spectrum (grid search), D = 4096, B = 1024
  6  the same signals with fnft_nsep's default localization (MIXED)
  7  config 3's signals through fnft_nsev with its default options (SUBSAMPLE_AND_REFINE)
  8  fnft_nsev reflection coefficient with the slow discretizations CF4_3 / CF5_3 / CF6_4, D = M = 1024, B = 256

For each: throughput of the batched C-ABI call with host buffers (wall clock around the
call, best of --reps), the reference library (oracle/_ref) on a bounded sample over all
host cores, and parity of the GPU results against that sample.  One JSON line per config.

    python scripts/bench_configs.py [--configs 3,4,5] [--scale 1.0] [--reps 3]
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------ signal generators
class InvOpts(C.Structure):
    _fields_ = [("discretization", C.c_int), ("contspec_type", C.c_int),
                ("contspec_inversion_method", C.c_int), ("discspec_type", C.c_int),
                ("max_iter", C.c_size_t), ("oversampling_factor", C.c_size_t)]


def _soliton_worker(args):
    from oracle import ref_lib as R
    lam, D, T = args
    L = R.lib()
    L.fnft_nsev_inverse_default_opts.restype = InvOpts
    o = L.fnft_nsev_inverse_default_opts()
    o.discspec_type = 0  # norming constants
    K = len(lam)
    rng = np.random.default_rng(int(abs(lam[0].real) * 1e6) % (2 ** 31))
    b = np.exp(1j * rng.uniform(0, 2 * np.pi, K))
    q = np.zeros(D, dtype=np.complex128)
    Ta = np.array(T, dtype=np.float64)
    XI = np.zeros(2)
    lam = np.ascontiguousarray(lam, dtype=np.complex128)
    L.fnft_nsev_inverse.argtypes = None
    ret = L.fnft_nsev_inverse(C.c_size_t(0), None, XI.ctypes.data_as(C.c_void_p), C.c_size_t(K),
                              lam.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                              C.c_size_t(D), q.ctypes.data_as(C.c_void_p),
                              Ta.ctypes.data_as(C.c_void_p), C.c_int32(1), C.byref(o))
    assert ret == 0, ret
    return q


def config3_inputs(B, D=4096, K=8, T=(-20.0, 20.0), seed=4096):
    rng = np.random.default_rng(seed)
    lams = []
    for _ in range(B):
        while True:
            lam = rng.uniform(-2, 2, K) + 1j * rng.uniform(0.3, 2.3, K)
            d = np.abs(lam[:, None] - lam[None, :]) + 10 * np.eye(K)
            if d.min() >= 0.1:
                break
        lams.append(lam)
    with mp.Pool(cores()) as pool:
        Q = pool.map(_soliton_worker, [(l, D, T) for l in lams], chunksize=8)
    guesses = np.array(lams) + 0.01 * (rng.normal(size=(B, K)) + 1j * rng.normal(size=(B, K)))
    return np.array(Q), np.array(lams), guesses


def config4_inputs(B, D=8192, T=(-16.0, 15.0), seed=8192):
    rng = np.random.default_rng(seed)
    t = np.linspace(T[0], T[1], D)[None, :]
    A = rng.uniform(0.5, 3.2, (B, 1))
    t0 = rng.uniform(-2, 2, (B, 1))
    w = rng.uniform(0.7, 1.5, (B, 1))
    return (A / np.cosh((t - t0) / w) ** 2).astype(np.complex128)


def config5_inputs(B, D=4096, seed=40960):
    rng = np.random.default_rng(seed)
    t = (2 * np.pi / D * np.arange(D))[None, :]
    A = rng.uniform(0.5, 2.5, (B, 1))
    m = rng.integers(0, 5, (B, 1))
    k = rng.integers(1, 5, (B, 1))
    e = rng.uniform(0, 0.3, (B, 1))
    ph = rng.uniform(0, 2 * np.pi, (B, 1))
    return A * np.exp(1j * m * t) * (1 + e * np.cos(k * t + ph))


# ------------------------------------------------------------------ reference workers
def _ref3(args):
    from oracle import ref_lib as R
    q, g, T = args
    R.lib().fnft_errwarn_setprintf(None)
    o = R.nsev_default_opts()
    o.bound_state_localization = 1  # NEWTON
    o.discspec_type = 2             # BOTH
    t0 = time.perf_counter()
    ret, cs, K, bs, nc = R.nsev(q, T, 0, None, 1, o, K=len(g), bound_states=g, want_contspec=False)
    return time.perf_counter() - t0, ret, K, bs, nc


def _ref4(args):
    from oracle import ref_lib as R
    u, T, M, XI = args
    o = R.lib().fnft_kdvv_default_opts()
    o.discretization = 19  # kdv 4SPLIT4B
    t0 = time.perf_counter()
    ret, cs = R.kdvv(u, T, M, XI, o)
    return time.perf_counter() - t0, ret, cs


def _ref5(args):
    from oracle import ref_lib as R
    q, T = args
    R.lib().fnft_errwarn_setprintf(None)
    o = R.lib().fnft_nsep_default_opts()
    o.localization = 1  # GRIDSEARCH
    o.filtering = 1     # MANUAL
    o.bounding_box[0], o.bounding_box[1], o.bounding_box[2], o.bounding_box[3] = -10, 10, -10, 10
    o.discretization = 11  # 2SPLIT4B
    t0 = time.perf_counter()
    ret, main, aux = R.nsep(q, T, 1, o)
    return time.perf_counter() - t0, ret, main, aux


def _ref8(args):
    from oracle import ref_lib as R
    q, T, M, XI, disc = args
    R.lib().fnft_errwarn_setprintf(None)
    o = R.nsev_default_opts()
    o.discretization = disc
    o.bound_state_localization = 1
    t0 = time.perf_counter()
    ret, cs, *_ = R.nsev(q, T, M, XI, -1, o)
    return time.perf_counter() - t0, ret, cs


def run_pool(fn, tasks):
    nc = cores()
    with mp.Pool(nc) as pool:
        pool.map(fn, tasks[:min(nc, len(tasks))])  # warm-up
        t0 = time.perf_counter()
        res = pool.map(fn, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    return res, wall


def best_of(fn, reps):
    ts = []
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--scale", type=float, default=1.0, help="batch size multiplier")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ref-signals", type=int, default=0)
    args = ap.parse_args()
    todo = [int(c) for c in args.configs.split(",")]
    from oracle import ref_lib as R
    have_ref = R.available()
    nc = cores()
    # reference legs first (fork before CUDA is initialised)
    ref = {}
    inputs = {}
    if 3 in todo or 7 in todo:
        B = max(16, int(1024 * args.scale))
        T = (-20.0, 20.0)
        Q, lam, G = config3_inputs(B) if have_ref else (None, None, None)
        inputs[3] = (Q, lam, G, T)
        if have_ref and 3 in todo:
            n = args.ref_signals or min(B, 4 * nc)
            ref[3] = run_pool(_ref3, [(Q[i], G[i], T) for i in range(n)])
    if 4 in todo:
        B = max(16, int(2048 * args.scale))
        U = config4_inputs(B)
        inputs[4] = U
        if have_ref:
            n = args.ref_signals or min(B, 4 * nc)
            ref[4] = run_pool(_ref4, [(U[i], (-16.0, 15.0), 8192, (-3.55, 3.95)) for i in range(n)])
    if 5 in todo or 6 in todo:
        B = max(16, int(1024 * args.scale))
        Q5 = config5_inputs(B)
        inputs[5] = Q5
        if have_ref and 5 in todo:
            n = args.ref_signals or min(B, nc)
            ref[5] = run_pool(_ref5, [(Q5[i], (0.0, 2 * np.pi)) for i in range(n)])

    if 8 in todo:
        B, D = max(int(256 * args.scale), 16), 1024
        T8, XI8 = (-16.0, 16.0), (-6.0, 6.0)
        t = np.linspace(T8[0], T8[1], D)
        rng = np.random.default_rng(1024)
        Q8 = np.stack([(0.5 + 2.0 * rng.random()) / np.cosh(t - rng.normal(0, 0.5)) *
                       np.exp(1j * rng.normal(0, 0.4) * t) for _ in range(B)])
        inputs[8] = Q8
        if have_ref:
            n = args.ref_signals or min(B, 2 * nc)
            for disc in (23, 24, 25):
                ref[(8, disc)] = run_pool(_ref8, [(Q8[i], T8, D, XI8, disc) for i in range(n)])

    import fnft_b200 as F
    F.lib().fnft_errwarn_setprintf(None)
    for cfg in todo:
        line = {"config": cfg, "n_gpus": 1, "cores": nc}
        if cfg == 3:
            Q, lam, G, T = inputs[3]
            if Q is None:
                print(json.dumps({"config": 3, "unavailable": "oracle/_ref needed to synthesise the solitons"}))
                continue
            B, K = G.shape
            o = F.nsev_default_opts()
            o.bound_state_localization = F.BSLOC_NEWTON
            o.discspec_type = F.DSTYPE_BOTH

            def run():
                return F.nsev_batch(Q, T, 0, None, 1, o, K=np.full(B, K), Kmax=K, bound_states=G)
            run()
            dt, (ret, cs, Ka, bs, ncs, rcs) = best_of(run, args.reps)
            line.update(workload="fnft_nsev Newton bound states + norming constants/residues, D=4096, "
                                 "8-soliton signals, B=%d" % B,
                        value=B / dt, unit="signals/s", ms_per_call=dt * 1e3, ret=int(ret),
                        found_all=float((Ka == K).mean()))
            if 3 in ref:
                res, wall = ref[3]
                line["cpu_baseline"] = {"value": len(res) / wall, "unit": "signals/s", "cores": nc,
                                        "kind": "reference", "sample": "%d signals" % len(res)}
                eb, en = 0.0, 0.0
                for i, (t, r, Kr, bsr, ncr) in enumerate(res):
                    if Kr != Ka[i]:
                        eb = np.inf
                        continue
                    for j in range(Kr):  # match by nearest eigenvalue (nsev_compare_nfs)
                        jj = int(np.argmin(np.abs(bs[i, :Kr] - bsr[j])))
                        eb = max(eb, abs(bs[i, jj] - bsr[j]) / abs(bsr[j]))
                        en = max(en, abs(ncs[i, jj] - ncr[j]) / abs(ncr[j]))
                        en = max(en, abs(ncs[i, K + jj] - ncr[Kr + j]) / abs(ncr[Kr + j]))
                line["parity"] = {"max_rel_err_bound_states": float(eb), "max_rel_err_normconsts_residues": float(en),
                                  "bound": 1e-9, "signals": len(res)}
        elif cfg == 4:
            U = inputs[4]
            B = U.shape[0]
            T, XI, M = (-16.0, 15.0), (-3.55, 3.95), 8192
            o = F.kdvv_default_opts()
            o.discretization = F.KDV_4SPLIT4B

            # pinned host buffers (as a production caller would use) when torch is available
            try:
                import torch
                Uh = torch.empty((B, U.shape[1]), dtype=torch.complex128, pin_memory=True)
                Uh.copy_(torch.from_numpy(U))
                csh = torch.empty((B, M), dtype=torch.complex128, pin_memory=True)
                rch = np.zeros(B, dtype=np.int32)
                Ta, XIa = np.array(T), np.array(XI)
                Lb = F.lib()

                def run():
                    r = Lb.fnft_kdvv_batch(B, U.shape[1], Uh.data_ptr(), Ta.ctypes.data, M, csh.data_ptr(),
                                           XIa.ctypes.data, C.addressof(o), rch.ctypes.data)
                    return r, csh.numpy(), rch
                line["host_buffers"] = "pinned"
            except ImportError:
                def run():
                    return F.kdvv_batch(U, T, M, XI, o)
                line["host_buffers"] = "pageable"
            run()
            dt, (ret, cs, rcs) = best_of(run, args.reps)
            line.update(workload="fnft_kdvv reflection coefficient, 4SPLIT4B, D=M=8192, B=%d" % B,
                        value=B / dt, unit="signals/s", ms_per_call=dt * 1e3, ret=int(ret))
            if 4 in ref:
                from common import parity_contract
                res, wall = ref[4]
                line["cpu_baseline"] = {"value": len(res) / wall, "unit": "signals/s", "cores": nc,
                                        "kind": "reference", "sample": "%d signals" % len(res)}
                worst = max(max(parity_contract(cs[i], res[i][2])) for i in range(len(res)))
                line["parity"] = {"parity_contract_max (must be < 1)": float(worst), "signals": len(res)}
        elif cfg == 5:
            Q5 = inputs[5]
            B, D = Q5.shape
            T = (0.0, 2 * np.pi)
            o = F.nsep_default_opts()
            o.localization = 1
            o.filtering = 1
            o.bounding_box[0], o.bounding_box[1], o.bounding_box[2], o.bounding_box[3] = -10, 10, -10, 10
            o.discretization = F.NSE_2SPLIT4B
            Kmax = Mmax = 2 * 2 * D
            out5 = F.nsep_buffers(B, Kmax, Mmax)  # allocated and touched once (see fnft_b200.nsep_buffers)

            def run():
                return F.nsep_batch(Q5, T, Kmax, Mmax, 1, o, out=out5)
            run()
            dt, (ret, Ka, main, Ma, aux, rcs) = best_of(run, args.reps)
            line.update(workload="fnft_nsep grid search (main + auxiliary spectrum), 2SPLIT4B, D=4096, B=%d" % B,
                        value=B / dt, unit="signals/s", ms_per_call=dt * 1e3, ret=int(ret))
            if 5 in ref:
                res, wall = ref[5]
                line["cpu_baseline"] = {"value": len(res) / wall, "unit": "signals/s", "cores": nc,
                                        "kind": "reference", "sample": "%d signals" % len(res)}
                ok, err = True, 0.0
                for i, (t, r, m0, a0) in enumerate(res):
                    m1, a1 = main[i, :int(Ka[i])], aux[i, :int(Ma[i])]
                    if len(m0) != len(m1) or len(a0) != len(a1):
                        ok = False
                        continue
                    if len(m0):
                        err = max(err, float(np.abs(m1 - m0).max() / max(1.0, np.abs(m0).max())))
                    if len(a0):
                        err = max(err, float(np.abs(a1 - a0).max() / max(1.0, np.abs(a0).max())))
                line["parity"] = {"same_counts": ok, "max_err": err, "bound": 1e-9, "signals": len(res)}
        elif cfg == 6:
            # config 5's signals with the reference's DEFAULT localization (MIXED: subsample-and-refine for
            # the non-real points + grid search for the real ones); no CPU leg: the oracle's stand-in for
            # eiscor is O(n^3) (minutes per signal at this size)
            Q5 = inputs[5]
            B, D = Q5.shape
            T = (0.0, 2 * np.pi)
            o = F.nsep_default_opts()
            o.filtering = 1
            o.bounding_box[0], o.bounding_box[1], o.bounding_box[2], o.bounding_box[3] = -10, 10, -10, 10
            o.discretization = F.NSE_2SPLIT4B
            Kmax = Mmax = 4 * 2 * D
            out6 = F.nsep_buffers(B, Kmax, Mmax)

            def run():
                return F.nsep_batch(Q5, T, Kmax, Mmax, 1, o, out=out6)
            run()
            dt, (ret, Ka, main, Ma, aux, rcs) = best_of(run, args.reps)
            line.update(workload="fnft_nsep default localization MIXED (main + auxiliary spectrum), 2SPLIT4B, "
                                 "D=4096, B=%d" % B,
                        value=B / dt, unit="signals/s", ms_per_call=dt * 1e3, ret=int(ret),
                        mean_main_points=float(Ka.mean()), mean_aux_points=float(Ma.mean()))
        elif cfg == 7:
            # config 3's multi-soliton signals through fnft_nsev with its DEFAULT options
            # (bsloc_SUBSAMPLE_AND_REFINE: GPU root finder on the subsampled signal + Newton), plus the
            # reflection coefficient at M = D points
            Q, lam, G, T = inputs[3]
            if Q is None:
                print(json.dumps({"config": 7, "unavailable": "oracle/_ref needed to synthesise the solitons"}))
                continue
            B, K = G.shape
            D = Q.shape[1]
            o = F.nsev_default_opts()
            o.discspec_type = F.DSTYPE_BOTH
            Kmax = 64
            # pinned host buffers, allocated once (as a production caller would: a fresh 64 MB np.zeros array for the
            # continuous spectrum costs its page faults inside every timed call, a pageable source 5 ms)
            Qh, cs7 = Q, None
            try:
                import torch
                Qt = torch.empty((B, D), dtype=torch.complex128, pin_memory=True)
                Qh = Qt.numpy()
                Qh[...] = Q
                cst = torch.empty((B, D), dtype=torch.complex128, pin_memory=True)
                cs7 = cst.numpy()
                cs7[...] = 0
                line["host_buffers"] = "pinned"
            except Exception:
                cs7 = np.zeros((B, D), dtype=np.complex128)
                line["host_buffers"] = "pageable"
            K0, G0 = np.zeros(B), np.zeros((B, Kmax), dtype=np.complex128)

            def run():
                return F.nsev_batch(Qh, T, D, (-4.0, 4.0), 1, o, K=K0, Kmax=Kmax, bound_states=G0, contspec_out=cs7)
            run()
            dt, (ret, cs, Ka, bs, ncs, rcs) = best_of(run, args.reps)
            ok = 0
            for i in range(B):  # every true eigenvalue found (BO-discretised, within 1e-3)
                ok += all(np.abs(bs[i, :int(Ka[i])] - l).min() < 2e-3 for l in lam[i]) if Ka[i] else 0
            line.update(workload="fnft_nsev default options (SUBSAMPLE_AND_REFINE bound states + residues + "
                                 "reflection coefficient, M=D), D=4096, 8-soliton signals, B=%d" % B,
                        value=B / dt, unit="signals/s", ms_per_call=dt * 1e3, ret=int(ret),
                        mean_K=float(Ka.mean()), all_eigenvalues_found=ok / B)
        elif cfg == 8:
            # slow (O(D*M)) commutator-free schemes: one warp per spectral point, D products of 2x2 exponentials
            Q, T, XI = inputs[8], T8, XI8
            B, D = Q.shape
            per = {}
            for name, disc in (("CF4_3", 23), ("CF5_3", 24), ("CF6_4", 25)):
                o = F.nsev_default_opts()
                o.discretization = disc
                o.bound_state_localization = 1

                def run():
                    return F.nsev_batch(Q, T, D, XI, -1, o)
                run()
                dt, out = best_of(run, args.reps)
                ret, cs = out[0], out[1]
                entry = {"signals_per_s": B / dt, "ms_per_call": dt * 1e3, "ret": int(ret)}
                if (8, disc) in ref:
                    res, wall = ref[(8, disc)]
                    nref = len(res)
                    entry["reference_signals_per_s"] = nref / wall
                    entry["max_rel_err_vs_reference"] = max(
                        float(np.abs(cs[i] - res[i][2]).sum() / np.abs(res[i][2]).sum()) for i in range(nref))
                per[name] = entry
            line.update(workload="fnft_nsev reflection coefficient, slow discretizations CF4_3 / CF5_3 / CF6_4, "
                                 "D=M=1024, kappa=-1, B=%d" % B, unit="signals/s", value=per["CF4_3"]["signals_per_s"],
                        per_scheme=per, cores=cores())
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

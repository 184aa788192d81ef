// fnft_b200 -- bound states with ONE EIGENVALUE PER WARP.
//
// Same mathematics as bound_kernels.cuh (fnft__nse_scatter_bound_states.c:281-338,
// :480-530, :639-654; Newton loop src/fnft_nsev.c:1007-1034), different work decomposition:
// the recurrence  [phi; dphi] <- [[U, 0], [U', U]] [phi; dphi]  is a product of block
// triangular matrices, so the D samples are split into 32 contiguous chunks, lane l
// multiplies up the 2x2 pair (P_l, P_l') of its chunk, and the 32 partial products are
// combined with warp shuffles (ordered tree reduction for Newton, ordered scan for the
// norming constants, where every lane then re-sweeps its chunk with the true start vector).
// 32x more parallelism per eigenvalue than one thread per eigenvalue, which is what fills
// the GPU at the batch sizes of BASELINE config 3 (1024 signals x 8 eigenvalues).
#pragma once
#ifndef FNFTB_EMUL
#include "bound_kernels.cuh"

struct BoMat {
    cplx m[4];  // product of the step matrices of a chunk
    cplx d[4];  // its derivative with respect to the spectral parameter
};

DEV void bo_mm(const cplx *A, const cplx *B, cplx *C)
{
    C[0] = cmul(A[0], B[0]);
    cfma(C[0], A[1], B[2]);
    C[1] = cmul(A[0], B[1]);
    cfma(C[1], A[1], B[3]);
    C[2] = cmul(A[2], B[0]);
    cfma(C[2], A[3], B[2]);
    C[3] = cmul(A[2], B[1]);
    cfma(C[3], A[3], B[3]);
}
DEV void bo_mm_acc(const cplx *A, const cplx *B, cplx *C)
{
    cfma(C[0], A[0], B[0]);
    cfma(C[0], A[1], B[2]);
    cfma(C[1], A[0], B[1]);
    cfma(C[1], A[1], B[3]);
    cfma(C[2], A[2], B[0]);
    cfma(C[2], A[3], B[2]);
    cfma(C[3], A[2], B[1]);
    cfma(C[3], A[3], B[3]);
}
// L after R:  (L.m * R.m,  L.d * R.m + L.m * R.d)
DEV BoMat bo_compose(const BoMat &L, const BoMat &R)
{
    BoMat o;
    bo_mm(L.m, R.m, o.m);
    bo_mm(L.d, R.m, o.d);
    bo_mm_acc(L.m, R.d, o.d);
    return o;
}
DEV cplx shfl_c(cplx v, int src)
{
    return make_cplx(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
DEV cplx shfl_down_c(cplx v, int off)
{
    return make_cplx(__shfl_down_sync(0xffffffffu, v.x, off), __shfl_down_sync(0xffffffffu, v.y, off));
}

// product (with derivative) of the forward steps of samples [lo, hi)
template <bool WITH_D>
DEV BoMat bo_chunk(const cplx *q, const cplx *r, int lo, int hi, cplx l, double h, bool descending, int wsel)
{
    BoMat P;
    P.m[0] = make_cplx(1.0, 0.0);
    P.m[1] = czero();
    P.m[2] = czero();
    P.m[3] = make_cplx(1.0, 0.0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        P.d[i] = czero();
    const int stp = (wsel >= FNFTB_WSEL_ES4) ? 3 : 1;  // ES4 / TES4: one step per grid point (3 samples)
    for (int k = lo; k < hi; k += stp) {
        const int n = descending ? (hi - stp - (k - lo)) : k;
        cplx U[4], Ud[4];
        slow_step_at<WITH_D>(q, r, n, l, h, wsel, -1.0, U, Ud);
        cplx t[4];
        if (WITH_D) {
            cplx td[4];
            bo_mm(Ud, P.m, td);
            bo_mm_acc(U, P.d, td);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                P.d[i] = td[i];
        }
        bo_mm(U, P.m, t);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            P.m[i] = t[i];
    }
    return P;
}

// chunk of lane `lane`: [lo, hi) in effective samples, aligned to the upsampling factor
DEV void bo_chunk_bounds(int D, int up, int lane, int *lo, int *hi)
{
    const int Dg = D / up;
    *lo = (int)(((long long)lane * Dg) / 32) * up;
    *hi = (int)(((long long)(lane + 1) * Dg) / 32) * up;
}

// Newton iterations, one warp per (signal, eigenvalue).  blockDim.x = 128.  The kernel is bound by the dependency chains
// of the step (sqrt, cosh/sinh, divisions: 49 % of the stalls are fixed-latency waits at 3 warps per scheduler, FP64 pipe
// 45 %, profiles/r02_roots_compact.md): four CTAs per SM (128 registers, 128 bytes of spills) beat three (166 registers),
// 6.99 -> 6.40 ms per 1024 signals of config 3; five (96 registers) measure the same as four.
__global__ void __launch_bounds__(128, 4) k_newton_warp(const BoundArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long gid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gid >= (long long)a.B * a.Kmax)
        return;
    const int s = (int)(gid / a.Kmax), i = (int)(gid % a.Kmax);
    if (i >= a.K[s])
        return;
    const cplx *q = a.q + (size_t)s * a.D;
    const cplx *rs = a.r ? a.r + (size_t)s * a.D : (const cplx *)0;
    int lo, hi;
    bo_chunk_bounds(a.D, a.upsampling, lane, &lo, &hi);
    cplx lam = a.lam[gid];
    const double im_max = a.box3 ? a.box3[s] : INFINITY;
    const double eprecision = 2.220446049250313e-16 * 100;
    const double tb = a.T0 - a.eps_t * a.bc;
    int iter = 0, status = 0;
    while (true) {
        const cplx l = cscale(lam, a.lweight);
        BoMat P = bo_chunk<true>(q, rs, lo, hi, l, a.eps_t, false, a.wsel);
        // ordered tree reduction: lane j ends up with P_(j+2^k-1) ... P_j
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            BoMat Hn;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                Hn.m[e] = shfl_down_c(P.m[e], off);
                Hn.d[e] = shfl_down_c(P.d[e], off);
            }
            if ((lane & (2 * off - 1)) == 0)
                P = bo_compose(Hn, P);
        }
        int go = 0;
        if (lane == 0) {
            // PHI(T0) = (exp(-i*lam*tb), 0), d/dlam = (-i*tb) * PHI
            const cplx p0 = c_exp(make_cplx(lam.y * tb, -lam.x * tb));
            const cplx d0 = cmul(p0, make_cplx(0.0, -tb));
            cplx phi[2], dphi[2], av, apv;
            phi[0] = cmul(P.m[0], p0);
            phi[1] = cmul(P.m[2], p0);
            dphi[0] = cmul(P.d[0], p0);
            cfma(dphi[0], P.m[0], d0);
            dphi[1] = cmul(P.d[2], p0);
            cfma(dphi[1], P.m[2], d0);
            bound_a_aprime(a, lam, phi, dphi, &av, &apv);
            if (av.x == 0.0 && av.y == 0.0) {
                go = 0;
            } else if (apv.x == 0.0 && apv.y == 0.0) {
                status = 3;
                go = 0;
            } else {
                const cplx err = cdiv(av, apv);
                lam = csub(lam, err);
                ++iter;
                go = 1;
                if (lam.y > im_max || lam.x > a.box1 || lam.x < a.box0 || lam.y < a.box2)
                    go = 0;
                else if (!(hypot(err.x, err.y) > eprecision && iter < a.niter))
                    go = 0;
            }
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        lam = shfl_c(lam, 0);
        if (!go)
            break;
    }
    if (lane == 0) {
        a.lam[gid] = lam;
        a.flag[gid] = status;
    }
}

// a, a', b for given eigenvalues, one warp per (signal, eigenvalue).  Scratch a.phi holds
// PHI at the given sample points of every eigenvalue: [koff[s] + i][D_given + 1][2].
// (launch bounds: five CTAs per SM, 96 registers: 5.60 -> 4.74 ms per 1024 signals of config 3; four: 5.12)
__global__ void __launch_bounds__(128, 5) k_normconsts_warp(const BoundArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long gid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gid >= (long long)a.B * a.Kmax)
        return;
    const int s = (int)(gid / a.Kmax), i = (int)(gid % a.Kmax);
    if (i >= a.K[s])
        return;
    const cplx *q = a.q + (size_t)s * a.D;
    const cplx *rs = a.r ? a.r + (size_t)s * a.D : (const cplx *)0;
    const int up = a.upsampling;
    const int Dg = a.D / up;
    int lo, hi;
    bo_chunk_bounds(a.D, up, lane, &lo, &hi);
    const cplx lcur = a.lam[gid];
    const cplx l = cscale(lcur, a.lweight);
    const double tb = a.T0 - a.eps_t * a.bc;
    const double te = a.T1 + a.eps_t * a.bc;
    cplx *store = a.phi + (size_t)(a.koff[s] + i) * (size_t)(Dg + 1) * 2;

    // ---- forward: chunk products, ordered scan for the start vectors ------------------
    {
        const BoMat P = bo_chunk<true>(q, rs, lo, hi, l, a.eps_t, false, a.wsel);
        cplx v1 = c_exp(make_cplx(lcur.y * tb, -lcur.x * tb)), v2 = czero();
        cplx w1 = cmul(v1, make_cplx(0.0, -tb)), w2 = czero();
        cplx s1 = v1, s2 = v2;  // start vector of this lane's chunk
        for (int j = 0; j < 32; ++j) {
            if (lane == j) {
                s1 = v1;
                s2 = v2;
            }
            cplx m[4], d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                m[e] = shfl_c(P.m[e], j);
                d[e] = shfl_c(P.d[e], j);
            }
            cplx nw1 = cmul(d[0], v1);
            cfma(nw1, d[1], v2);
            cfma(nw1, m[0], w1);
            cfma(nw1, m[1], w2);
            cplx nw2 = cmul(d[2], v1);
            cfma(nw2, d[3], v2);
            cfma(nw2, m[2], w1);
            cfma(nw2, m[3], w2);
            cplx nv1 = cmul(m[0], v1);
            cfma(nv1, m[1], v2);
            cplx nv2 = cmul(m[2], v1);
            cfma(nv2, m[3], v2);
            v1 = nv1;
            v2 = nv2;
            w1 = nw1;
            w2 = nw2;
        }
        if (lane == 0) {
            cplx phi[2] = {v1, v2}, dphi[2] = {w1, w2}, av, apv;
            bound_a_aprime(a, lcur, phi, dphi, &av, &apv);
            a.a_out[gid] = av;
            a.ap_out[gid] = apv;
            store[0] = s1;  // PHI at the first sample point (lane 0 starts there)
            store[1] = s2;
        }
        // re-sweep the chunk with the true start vector, storing PHI at the given points
        cplx p1 = s1, p2 = s2;
        for (int n = lo; n < hi;) {
            cplx U[4], Ud[4];
            n += slow_step_at<false>(q, rs, n, l, a.eps_t, a.wsel, -1.0, U, Ud);
            cplx g = cmul(U[0], p1);
            cfma(g, U[1], p2);
            cplx f = cmul(U[2], p1);
            cfma(f, U[3], p2);
            p1 = g;
            p2 = f;
            if ((n % up) == 0) {  // n = first sample after this step
                const size_t ng = (size_t)n / up;
                store[ng * 2] = p1;
                store[ng * 2 + 1] = p2;
            }
        }
    }
    __syncwarp();
    // ---- backward: chunk products (descending n), ordered scan from the last lane ------
    double best = INFINITY;
    int best_n = 0x7fffffff;
    cplx bval = czero();
    {
        const BoMat P = bo_chunk<false>(q, rs, lo, hi, l, -a.eps_t, true, a.wsel);
        cplx v1 = czero(), v2 = c_exp(make_cplx(-lcur.y * te, lcur.x * te));
        cplx s1 = v1, s2 = v2;
        for (int j = 31; j >= 0; --j) {
            if (lane == j) {
                s1 = v1;
                s2 = v2;
            }
            cplx m[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                m[e] = shfl_c(P.m[e], j);
            cplx nv1 = cmul(m[0], v1);
            cfma(nv1, m[1], v2);
            cplx nv2 = cmul(m[2], v1);
            cfma(nv2, m[3], v2);
            v1 = nv1;
            v2 = nv2;
        }
        cplx psi1 = s1, psi2 = s2;
        const int stp = (a.wsel >= FNFTB_WSEL_ES4) ? 3 : 1;
        for (int n = hi - stp; n >= lo; n -= stp) {
            cplx U[4], Ud[4];
            slow_step_at<false>(q, rs, n, l, -a.eps_t, a.wsel, -1.0, U, Ud);
            cplx d = cmul(U[0], psi1);
            cfma(d, U[1], psi2);
            cplx c = cmul(U[2], psi1);
            cfma(c, U[3], psi2);
            psi1 = d;
            psi2 = c;
            if ((n % up) == 0) {
                const int ng = n / up;
                const cplx p1 = store[(size_t)ng * 2], p2 = store[(size_t)ng * 2 + 1];
                // tmp = |0.5*log(|(PHI2/PSI2)/(PHI1/PSI1)|)|  (:642-654)
                const cplx r2 = cdiv(p2, psi2), r1 = cdiv(p1, psi1);
                const cplx rr = cdiv(r2, r1);
                const double tmp = fabs(0.5 * log(hypot(rr.x, rr.y)));
                // the reference scans ascending and keeps the first strict minimum
                if (tmp <= best) {
                    best = tmp;
                    best_n = ng;
                    bval = r1;
                }
            }
        }
    }
    // smallest metric, ties -> smallest sample index; NaN metrics never win
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int on = __shfl_xor_sync(0xffffffffu, best_n, off);
        const cplx ov = make_cplx(__shfl_xor_sync(0xffffffffu, bval.x, off),
                                  __shfl_xor_sync(0xffffffffu, bval.y, off));
        if (ob < best || (ob == best && on < best_n)) {
            best = ob;
            best_n = on;
            bval = ov;
        }
    }
    if (lane == 0)
        a.b_out[gid] = bval;
}
#endif  // !FNFTB_EMUL

// fnft_b200 -- all roots of a batch of complex polynomials (Aberth-Ehrlich iteration).
//
// Replaces fnft__poly_roots_fasteigen
//   /root/reference/src/private/fnft__poly_roots_fasteigen.c:29-48
// i.e. the eiscor companion-pencil QR solver (Fortran, z_poly_roots_modified_) that the
// reference uses for fnft_nsev_bsloc_FAST_EIGENVALUE / _SUBSAMPLE_AND_REFINE
// (src/fnft_nsev.c:687-711, 272-303).  A sequential unitary-plus-rank-one QR does not map
// to a GPU; simultaneous iteration does: every root z_i is updated by
//
//     N_i = p(z_i)/p'(z_i),   S_i = sum_{j != i} 1/(z_i - z_j),   z_i <- z_i - N_i/(1 - N_i S_i)
//
// (cubically convergent, embarrassingly parallel over i, O(n^2) per sweep).
//
//   * one CTA per polynomial, roots in shared memory (degree <= 8192; in the output array in global
//     memory up to 32768), groups of roots updated Jacobi style, Gauss-Seidel from group to group, coefficients read through the read-only path (all lanes
//     read the same address: broadcast);
//   * p and p' by Horner's rule in double; for |z| > 1 on the reversed polynomial in w = 1/z
//     (p(z) = z^n q(w),  N = z/(n - w q'(w)/q(w))), so nothing overflows;
//   * S_i only steers the convergence -- the fixed points are exactly the zeros of N_i -- so
//     it is accumulated in single precision (differences formed in double, one MUFU
//     reciprocal per term): the O(n^2) part costs 1/20 of what double division would;
//   * a root is frozen once |p(z)| <= 4 n eps sum |c_k||z|^k (backward-stable root of a
//     polynomial whose coefficients are perturbed by a few ulp -- the accuracy eiscor's
//     backward-stable QR delivers norm-wise, here coefficient-wise);
//   * start values: Bini's rule -- moduli from the upper convex hull of (k, log|a_k|), equally
//     spaced arguments on every circle, golden-ratio offsets between the circles;
//   * values that are still moving after maxit sweeps are returned as NaN (the callers' box
//     filter drops them; fnft__poly_roots_fasteigen reports the failure like eiscor's info).
//   * degree <= 8192 (round 2): k_roots_aberth_c below -- the sweeps walk an ordered, compacted list of the roots that
//     still move, the Aberth sums run in packed single precision on float hi/lo copies of the roots
//     (profiles/r02_roots_compact.md: 3.4x faster on BASELINE config 7); k_roots_aberth is the round-1 kernel that
//     still serves degrees 8193 ... 32768 and FNFT_B200_ROOTS_COMPACT=0.
// Exactly zero leading / trailing coefficients are split off (roots reported as 0, which
// the callers' z -> lambda map sends to infinity and the bounding-box filter drops).
#pragma once
#include "launch.cuh"

#ifndef FNFTB_EMUL

struct RootsArgs {
    const cplx *coef;   // polynomial b at coef + b*cstride: n+1 coefficients, highest power first
    long long cstride;
    int n;              // nominal degree
    cplx *roots;        // [B][n]
    double *absc;       // [B][n+1] workspace: |c_k|
    double *lg;         // [B][n+1] workspace: log|a_i| of the stripped polynomial, ascending powers
    int *hull;          // [B][n+2] workspace: vertices of the upper convex hull
    int *info;          // [B][4]: lead, m (effective degree), sweeps used, roots not converged
    int maxit;
    int in_global;      // 1: the sweeps work on `roots` in global memory (degree > 8192)
    int stats;          // 1: the compacting kernel adds its work counters to g_roots_stat
};

// ---- start values -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_roots_init(const RootsArgs a)
{
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int n = a.n;
    const cplx *c = a.coef + (size_t)b * a.cstride;
    double *absc = a.absc + (size_t)b * (n + 1);
    double *lg = a.lg + (size_t)b * (n + 1);
    int *hull = a.hull + (size_t)b * (n + 2);
    cplx *roots = a.roots + (size_t)b * n;
    __shared__ int s_lead, s_last, s_nh;
    if (tid == 0) {
        s_lead = n + 1;
        s_last = -1;
    }
    __syncthreads();
    int lead = n + 1, last = -1;
    for (int k = tid; k <= n; k += nt) {
        const cplx v = c[k];
        const double m = hypot(v.x, v.y);
        absc[k] = m;
        if (m > 0.0 && m < INFINITY) {
            lead = min(lead, k);
            last = max(last, k);
        }
    }
    atomicMin(&s_lead, lead);
    atomicMax(&s_last, last);
    __syncthreads();
    lead = s_lead;
    last = s_last;
    const int m = (last >= lead) ? last - lead : 0;  // effective degree after stripping zeros
    // a_i (ascending powers, i = 0..m) = c[last - i]
    for (int i = tid; i <= m; i += nt)
        lg[i] = (last >= lead && absc[last - i] > 0.0) ? log(absc[last - i]) : -INFINITY;
    for (int i = tid; i < n; i += nt)
        roots[i] = czero();
    __syncthreads();
    if (tid == 0) {
        // upper convex hull of (i, lg[i]) by a monotone chain
        int nh = 0;
        for (int i = 0; i <= m; ++i) {
            if (lg[i] == -INFINITY)
                continue;
            while (nh >= 2) {
                const int i1 = hull[nh - 2], i2 = hull[nh - 1];
                // drop i2 if it lies on or below the chord i1 -> i
                const double cross = (double)(i2 - i1) * (lg[i] - lg[i1]) - (double)(i - i1) * (lg[i2] - lg[i1]);
                if (cross >= 0.0)
                    --nh;
                else
                    break;
            }
            hull[nh++] = i;
        }
        s_nh = nh;
        a.info[4 * b + 0] = lead;
        a.info[4 * b + 1] = m;
        a.info[4 * b + 2] = 0;
        a.info[4 * b + 3] = 0;
    }
    __syncthreads();
    const int nh = s_nh;
    // roots lead .. lead+m-1 of the output belong to the stripped polynomial
    for (int s = 0; s + 1 < nh; ++s) {
        const int lo = hull[s], hi = hull[s + 1], cnt = hi - lo;
        const double r = exp((lg[lo] - lg[hi]) / (double)cnt);
        for (int j = tid; j < cnt; j += nt) {
            // Bini's rule spaces the cnt roots of a hull segment equally; the offset between
            // segments is a golden-ratio sequence here, so that many short segments (a(z) of a
            // pulse has hundreds of 2-root segments) still cover all arguments evenly -- with
            // Bini's offset 2 pi s/m they start in two narrow sectors and crawl along the ring
            const double off = (double)s * 0.6180339887498949;
            const double th = 2.0 * 3.141592653589793 * (((double)j + (off - floor(off))) / (double)cnt) + 0.7;
            double sn, cs;
            sincos(th, &sn, &cs);
            roots[lo + j] = make_cplx(r * cs, r * sn);
        }
    }
}

// ---- Aberth-Ehrlich sweeps -----------------------------------------------------------------
// Every thread owns G groups of R roots (root (g*R + r)*nt + tid).  The R roots of a group are
// processed together: one pass over the coefficients serves their R Horner recurrences (forward
// coefficient order for |z| <= 1, reversed for |z| > 1 -- both are read, the recurrence picks), one
// pass over the shared-memory copy of the roots serves their R Aberth sums.  A sweep is Jacobi
// within a group and Gauss-Seidel from group to group (barrier, publish, barrier).
template <int R, int MAXNT>
__global__ void __launch_bounds__(MAXNT) k_roots_aberth(const RootsArgs a)
{
    extern __shared__ double2 fnftb_smem[];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int n = a.n;
    const int lead = a.info[4 * b + 0], m = a.info[4 * b + 1];
    if (m <= 0)
        return;
    // stripped polynomial: coefficients c[lead .. lead+m], highest power first
    const cplx *c = a.coef + (size_t)b * a.cstride + lead;
    const double *ac = a.absc + (size_t)b * (n + 1) + lead;
    cplx *groots = a.roots + (size_t)b * n;
    // working copy of the roots: shared memory up to degree 8192, else the output array itself
    // (every access below is a broadcast read or a thread's own element; barriers order the sweeps)
    cplx *z = a.in_global ? groots : (cplx *)fnftb_smem;  // [m]
    if (!a.in_global) {
        for (int i = tid; i < m; i += nt)
            z[i] = groots[i];
    }
    __syncthreads();
    const int G = (m + R * nt - 1) / (R * nt);  // <= 32 / R
    unsigned done = 0;  // bit g*R + r: that root has converged (or does not exist)
    for (int q = 0; q < G * R; ++q)
        if (tid + q * nt >= m)
            done |= (1u << q);
    const unsigned all_done = (G * R >= 32) ? 0xffffffffu : ((1u << (G * R)) - 1u);
    const double tol = 4.0 * (double)m * 2.220446049250313e-16;
    int it = 0, left = 1;
    for (; it < a.maxit && left > 0; ++it) {
        for (int g = 0; g < G; ++g) {
            const unsigned gdone = (done >> (g * R)) & ((1u << R) - 1u);
            unsigned ndone = gdone;
            cplx zi[R], w[R], p[R], dp[R];  // p is reused for the Newton correction, zi for the new value
            double aw[R], e[R];
            bool small[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = tid + (g * R + r) * nt;
                zi[r] = (i < m) ? z[i] : make_cplx(0.5, 0.0);
                const double az = hypot(zi[r].x, zi[r].y);
                small[r] = (az <= 1.0);
                w[r] = small[r] ? zi[r] : cdiv(make_cplx(1.0, 0.0), zi[r]);
                aw[r] = small[r] ? az : 1.0 / az;
                p[r] = czero();
                dp[r] = czero();
                e[r] = 0.0;
            }
            if (gdone != (1u << R) - 1u) {
                // Horner for p, p' and the running error bound: p(z) for |z| <= 1, q(1/z) = p(z)/z^m else
                if constexpr (R == 1) {
                    // one coefficient stream, in the direction this root needs (same arithmetic as below)
                    const cplx *cp = small[0] ? c : c + m;
                    const double *ap = small[0] ? ac : ac + m;
                    const int stp = small[0] ? 1 : -1;
                    for (int k = 0; k <= m; ++k) {
                        const cplx ck = LDG(cp + stp * k);
                        const double ak = LDG(ap + stp * k);
                        dp[0] = cadd(cmul(dp[0], w[0]), p[0]);
                        p[0] = cadd(cmul(p[0], w[0]), ck);
                        e[0] = e[0] * aw[0] + ak;
                    }
                } else {
                    for (int k = 0; k <= m; ++k) {
                        const cplx cf = LDG(&c[k]), cb = LDG(&c[m - k]);
                        const double af = LDG(&ac[k]), ab = LDG(&ac[m - k]);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const cplx ck = small[r] ? cf : cb;
                            dp[r] = cadd(cmul(dp[r], w[r]), p[r]);
                            p[r] = cadd(cmul(p[r], w[r]), ck);
                            e[r] = e[r] * aw[r] + (small[r] ? af : ab);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((ndone >> r) & 1u)
                        continue;
                    if (hypot(p[r].x, p[r].y) <= tol * e[r]) {  // frozen from now on
                        ndone |= (1u << r);
                        continue;
                    }
                    cplx Nw;
                    if (small[r]) {
                        Nw = cdiv(p[r], dp[r]);
                    } else {  // N = z / (m - w q'(w)/q(w))
                        const cplx t = cmul(w[r], cdiv(dp[r], p[r]));
                        Nw = cdiv(zi[r], make_cplx((double)m - t.x, -t.y));
                    }
                    if (!(isfinite(Nw.x) && isfinite(Nw.y))) {  // p' = 0: leave the stationary point sideways
                        const double az = hypot(zi[r].x, zi[r].y);
                        Nw = make_cplx(1e-3 * az + 1e-6, 1e-3 * az + 1e-6);
                    }
                    p[r] = Nw;
                }
                if (ndone != (1u << R) - 1u) {
                    float sx[R], sy[R];
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        sx[r] = sy[r] = 0.f;
                    for (int j = 0; j < m; ++j) {
                        const cplx zj = z[j];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float dx = (float)(zi[r].x - zj.x), dy = (float)(zi[r].y - zj.y);
                            const float r2 = fmaxf(dx * dx + dy * dy, 1e-37f);
                            const float inv = __frcp_rn(r2);
                            sx[r] = fmaf(dx, inv, sx[r]);
                            sy[r] = fmaf(-dy, inv, sy[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if ((ndone >> r) & 1u)
                            continue;
                        const cplx S = make_cplx((double)sx[r], (double)sy[r]);
                        const cplx den = csub(make_cplx(1.0, 0.0), cmul(p[r], S));
                        cplx dz = cdiv(p[r], den);
                        if (!(isfinite(dz.x) && isfinite(dz.y)))
                            dz = p[r];
                        zi[r] = csub(zi[r], dz);
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (!((ndone >> r) & 1u))
                    z[tid + (g * R + r) * nt] = zi[r];
            done |= ndone << (g * R);
            __syncthreads();
        }
        left = __syncthreads_count(done != all_done);
    }
    int notconv = 0;
    for (int q = 0; q < G * R; ++q) {
        const int i = tid + q * nt;
        if (i < m) {
            if (!((done >> q) & 1u)) {  // still moving after maxit sweeps: not a root, say so
                ++notconv;
                groots[i] = make_cplx(nan(""), nan(""));
            } else if (!a.in_global) {
                groots[i] = z[i];
            }
        }
    }
    if (notconv)
        atomicAdd(&a.info[4 * b + 3], notconv);
    if (tid == 0)
        a.info[4 * b + 2] = it;
}

// ---- Aberth-Ehrlich sweeps with compaction of the roots that still move ------------------------
// The kernel above keeps a warp busy until the last of its 32 roots has converged.  Here the roots that
// still move are kept as an ordered list in shared memory: every sweep walks the list in chunks of
// blockDim.x roots (Jacobi inside a chunk, Gauss-Seidel from chunk to chunk), rebuilds it by an ordered
// compaction (ballots, warp counts, no atomics: the schedule and with it every bit of the result is
// reproducible), and the work of a sweep shrinks with the list.  The double-precision roots stay in the
// output array (a root is read and written once per sweep by the thread that has it); shared memory
// holds every root as a pair of floats per component (hi = (float)x, lo = (float)(x - hi)), from which
// the Aberth sum forms its differences as (hi_i - hi_j) + (lo_i - lo_j) -- exact to 2^-48 |z| for
// neighbouring roots and free of FP64 -> FP32 conversions, which issue at a quarter of the FP64 rate.
__device__ unsigned long long g_roots_stat[4];  // FNFT_B200_ROOTS_STATS=1: root-sweeps, polynomials, sweeps, roots

// one MUFU.RCP (__frcp_rn adds a Newton step, a range test and a branch to a slow path: twice the instructions of the loop)
static __device__ __forceinline__ float roots_rcp(const float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

static __device__ __forceinline__ float4 roots_split(const cplx v)
{
    const float hx = (float)v.x, hy = (float)v.y;
    return make_float4(hx, hy, (float)(v.x - (double)hx), (float)(v.y - (double)hy));
}

// CS: the coefficients and their moduli are staged in shared memory as well (44 instead of 20 bytes per root): the Horner
// recurrences of a CTA's last few roots then run at shared-memory latency instead of waiting for L2 every few steps.
// shared-memory copy of root i for the Aberth sums: slot i & 1 of the pair i >> 1, negated (the loop only adds)
static __device__ __forceinline__ void roots_publish(float4 *zf, const int i, const cplx v)
{
    const float4 f = roots_split(v);
    float *hi = (float *)(zf + 2 * (i >> 1)), *lo = hi + 4;
    const int sl = i & 1;
    hi[sl] = -f.x;
    hi[2 + sl] = -f.y;
    lo[sl] = -f.z;
    lo[2 + sl] = -f.w;
}

template <int MAXNT, bool CS>
__global__ void __launch_bounds__(MAXNT) k_roots_aberth_c(const RootsArgs a)
{
    extern __shared__ double2 fnftb_smem[];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
    const int n = a.n;
    const int lead = a.info[4 * b + 0], m = a.info[4 * b + 1];
    if (m <= 0)
        return;
    const cplx *c = a.coef + (size_t)b * a.cstride + lead;
    const double *ac = a.absc + (size_t)b * (n + 1) + lead;
    cplx *groots = a.roots + (size_t)b * n;
    // roots 2p and 2p+1 as NEGATED float pairs, packed for the f32x2 instructions of sm_100:
    // zf[2p] = -(hi.x0, hi.x1, hi.y0, hi.y1), zf[2p+1] = -(lo.x0, lo.x1, lo.y0, lo.y1)
    float4 *zf = (float4 *)fnftb_smem;                  // [n + 1]
    cplx *sc = (cplx *)(zf + n + 1);                    // [n + 1] (CS)
    double *sa = (double *)(sc + (CS ? n + 1 : 0));     // [n + 1] (CS)
    int *s_wcnt = (int *)(sa + (CS ? n + 1 : 0));       // [32]
    unsigned short *lst0 = (unsigned short *)(s_wcnt + 32);  // [n] roots that still move, ascending
    unsigned short *lst1 = lst0 + n;                         // [n] the list of the next sweep
    for (int i = tid; i < m; i += nt) {
        roots_publish(zf, i, groots[i]);
        lst0[i] = (unsigned short)i;
    }
    if (tid == 0 && (m & 1))  // partner of the last root of an odd degree: far away, its term is (-1e30) * rcp(inf) = 0
        roots_publish(zf, m, make_cplx(1e30, 0.0));
    if constexpr (CS) {
        for (int i = tid; i <= m; i += nt) {
            sc[i] = c[i];
            sa[i] = ac[i];
        }
    }
    __syncthreads();
    const double tol = 4.0 * (double)m * 2.220446049250313e-16;
    int it = 0, nact = m;
    unsigned long long work = 0;
    for (; it < a.maxit && nact > 0; ++it) {
        work += (unsigned long long)nact;
        const unsigned short *cur = (it & 1) ? lst1 : lst0;
        unsigned short *nxt = (it & 1) ? lst0 : lst1;
        int nnext = 0;
        for (int base = 0; base < nact; base += nt) {
            const int k = base + tid;
            bool moving = false;
            int i = 0;
            cplx zi = czero();
            if (k < nact) {
                i = cur[k];
                zi = groots[i];
                const double az = hypot(zi.x, zi.y);
                const bool small = (az <= 1.0);
                const cplx w = small ? zi : cdiv(make_cplx(1.0, 0.0), zi);
                const double aw = small ? az : 1.0 / az;
                cplx p = czero(), dp = czero();
                double e = 0.0;
                // Horner for p, p' and the running error bound: p(z) for |z| <= 1, q(1/z) = p(z)/z^m else
                const cplx *cp = small ? (CS ? sc : c) : (CS ? sc : c) + m;
                const double *ap = small ? (CS ? sa : ac) : (CS ? sa : ac) + m;
                const int stp = small ? 1 : -1;
#pragma unroll 4
                for (int q = 0; q <= m; ++q) {
                    cplx ck;
                    double ak;
                    if constexpr (CS) {
                        ck = cp[stp * q];
                        ak = ap[stp * q];
                    } else {
                        ck = LDG(cp + stp * q);
                        ak = LDG(ap + stp * q);
                    }
                    // dp = dp w + p, p = p w + c_k as four fused multiply-adds each
                    const double dx = fma(dp.x, w.x, fma(-dp.y, w.y, p.x));
                    dp.y = fma(dp.x, w.y, fma(dp.y, w.x, p.y));
                    dp.x = dx;
                    const double px = fma(p.x, w.x, fma(-p.y, w.y, ck.x));
                    p.y = fma(p.x, w.y, fma(p.y, w.x, ck.y));
                    p.x = px;
                    e = fma(e, aw, ak);
                }
                if (!(hypot(p.x, p.y) <= tol * e)) {  // else: converged, leaves the list
                    moving = true;
                    cplx Nw;
                    if (small) {
                        Nw = cdiv(p, dp);
                    } else {  // N = z / (m - w q'(w)/q(w))
                        const cplx t = cmul(w, cdiv(dp, p));
                        Nw = cdiv(zi, make_cplx((double)m - t.x, -t.y));
                    }
                    if (!(isfinite(Nw.x) && isfinite(Nw.y))) {  // p' = 0: leave the stationary point sideways
                        Nw = make_cplx(1e-3 * az + 1e-6, 1e-3 * az + 1e-6);
                    }
                    const float4 fi = roots_split(zi);
                    const float2 fhx = make_float2(fi.x, fi.x), fhy = make_float2(fi.y, fi.y);
                    const float2 flx = make_float2(fi.z, fi.z), fly = make_float2(fi.w, fi.w);
                    float2 sx = make_float2(0.f, 0.f), sy = make_float2(0.f, 0.f);
                    const int np = (m + 1) >> 1;
                    // two roots per iteration in packed single precision (FADD2 / FMUL2 / FFMA2): differences
                    // (hi_i - hi_j) + (lo_i - lo_j), 1 / |d|^2 by MUFU.RCP, sums of d / |d|^2 (conjugated at the end)
#pragma unroll 4
                    for (int pj = 0; pj < np; ++pj) {
                        const float4 A = zf[2 * pj], Bq = zf[2 * pj + 1];
                        const float2 dx = __fadd2_rn(__fadd2_rn(fhx, make_float2(A.x, A.y)), __fadd2_rn(flx, make_float2(Bq.x, Bq.y)));
                        const float2 dy = __fadd2_rn(__fadd2_rn(fhy, make_float2(A.z, A.w)), __fadd2_rn(fly, make_float2(Bq.z, Bq.w)));
                        const float2 r2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
                        const float2 inv = make_float2(roots_rcp(fmaxf(r2.x, 1e-37f)), roots_rcp(fmaxf(r2.y, 1e-37f)));
                        sx = __ffma2_rn(dx, inv, sx);
                        sy = __ffma2_rn(dy, inv, sy);
                    }
                    const float sx0 = sx.x, sx1 = sx.y, sy0 = -sy.x, sy1 = -sy.y;
                    const cplx S = make_cplx((double)sx0 + (double)sx1, (double)sy0 + (double)sy1);
                    const cplx den = csub(make_cplx(1.0, 0.0), cmul(Nw, S));
                    cplx dz = cdiv(Nw, den);
                    if (!(isfinite(dz.x) && isfinite(dz.y)))
                        dz = Nw;
                    zi = csub(zi, dz);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, moving);
            __syncthreads();  // every Aberth sum of this chunk has read the old values
            if (moving) {
                groots[i] = zi;
                roots_publish(zf, i, zi);
            }
            if (lane == 0)
                s_wcnt[wid] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int q = 0; q < nw; ++q) {
                const int cq = s_wcnt[q];
                before += (q < wid) ? cq : 0;
                total += cq;
            }
            if (moving)
                nxt[nnext + before + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)i;
            nnext += total;
            // s_wcnt is rewritten only after the first barrier of the next chunk, nxt is read only in the next sweep
        }
        __syncthreads();
        nact = nnext;
    }
    // still moving after maxit sweeps: not a root, say so
    const unsigned short *cur = (it & 1) ? lst1 : lst0;
    for (int k = tid; k < nact; k += nt)
        groots[cur[k]] = make_cplx(nan(""), nan(""));
    if (tid == 0) {
        a.info[4 * b + 2] = it;
        a.info[4 * b + 3] = nact;
        if (a.stats) {
            atomicAdd(&g_roots_stat[0], work);
            atomicAdd(&g_roots_stat[1], 1ull);
            atomicAdd(&g_roots_stat[2], (unsigned long long)it);
            atomicAdd(&g_roots_stat[3], (unsigned long long)m);
        }
    }
}

static inline int roots_launch_c(const RootsArgs &a, int B, cudaStream_t st)
{
    // CTA size: with many polynomials in flight small CTAs (6 - 8 per SM) fill the SM while others are down to
    // their last few roots; a single polynomial wants as many threads as it has roots
    static const int knob_cs = [] {
        const char *e = getenv("FNFT_B200_ROOTS_SMEM_COEF");
        return e ? atoi(e) : 1;
    }();
    static const int knob_nt = [] {
        const char *e = getenv("FNFT_B200_ROOTS_NT");
        return e ? atoi(e) : 0;
    }();
    const bool cs = knob_cs && a.n <= 4096;
    int want = (B >= 296) ? 256 : (B >= 74) ? 512 : 1024;
    if (knob_nt)
        want = knob_nt;
    const int nt = std::min(want, ((a.n + 31) / 32) * 32);
    const size_t smem = sizeof(float4) * ((size_t)a.n + 1) + 2 * sizeof(unsigned short) * (size_t)a.n + 32 * sizeof(int) +
                        (cs ? (sizeof(cplx) + sizeof(double)) * ((size_t)a.n + 1) : 0);
    auto kern = cs ? k_roots_aberth_c<1024, true> : k_roots_aberth_c<1024, false>;
    {
        const int e = fnftb_smem_optin((const void *)kern, smem);
        if (e != 0)
            return e;
    }
    kern<<<B, nt, smem, st>>>(a);
    return 0;
}

template <int R, int MAXNT>
static inline int roots_launch_r(const RootsArgs &a, int B, int nt, cudaStream_t st)
{
    const size_t smem = a.in_global ? 0 : sizeof(cplx) * (size_t)a.n;
    auto kern = k_roots_aberth<R, MAXNT>;
    {
        const int e = fnftb_smem_optin((const void *)kern, smem);
        if (e != 0)
            return e;
    }
    kern<<<B, nt, smem, st>>>(a);
    return 0;
}

// Launches both kernels for B polynomials of nominal degree n.  Returns 0, or -6 when n
// exceeds 32768 (roots in shared memory up to 8192, in global memory beyond).
static inline int roots_launch(const RootsArgs &a_in, int B, cudaStream_t st)
{
    if (a_in.n < 1)
        return -2;
    if (a_in.n > 32768)  // 32 roots per thread (one bit each in `done`)
        return -6;
    RootsArgs a = a_in;
    a.in_global = (a.n > 8192) ? 1 : 0;
    k_roots_init<<<B, 256, 0, st>>>(a);
    ++g_fnftb_launch_count;
    // Up to degree 8192 (the float copies of the roots fit in shared memory): the compacting kernel.  Config 7
    // (1024 polynomials of degree 1638 / 2048): see profiles/r02_roots_compact.md; FNFT_B200_ROOTS_COMPACT=0 selects the
    // kernel below.
    static const int knob_compact = [] {
        const char *e = getenv("FNFT_B200_ROOTS_COMPACT");
        return e ? atoi(e) : 1;
    }();
    if (knob_compact && a.n <= 8192) {
        static const int knob_stats = [] {
            const char *e = getenv("FNFT_B200_ROOTS_STATS");
            return e ? atoi(e) : 0;
        }();
        a.in_global = 1;
        a.stats = knob_stats;
        if (knob_stats) {
            const unsigned long long z4[4] = {0, 0, 0, 0};
            cudaMemcpyToSymbolAsync(g_roots_stat, z4, sizeof(z4), 0, cudaMemcpyHostToDevice, st);
        }
        const int rc = roots_launch_c(a, B, st);
        if (rc)
            return rc;
        ++g_fnftb_launch_count;
        if (knob_stats) {
            unsigned long long s4[4];
            cudaStreamSynchronize(st);
            cudaMemcpyFromSymbol(s4, g_roots_stat, sizeof(s4));
            fprintf(stderr, "[roots] %llu polynomials, %llu roots, %.2f sweeps per polynomial, %.2f sweeps per root\n", s4[1],
                    s4[3], (double)s4[2] / (double)(s4[1] ? s4[1] : 1), (double)s4[0] / (double)(s4[3] ? s4[3] : 1));
        }
        return (int)cudaGetLastError();
    }
    // One root per group (R = 1) measured fastest on B200: converged roots drop out one by one,
    // which saves more than sharing the coefficient / root loads between the R roots of a group
    // (R = 4 with 416 threads: 259 ms, R = 1 with 832 threads: 134 ms for 1024 polynomials of
    // degree 1638).  CTA size: every thread owns G = ceil(n / 1024) roots (up to rounding).
    const int G = (a.n + 1023) / 1024;
    const int nt = (((a.n + G - 1) / G + 31) / 32) * 32;
    const int rc = roots_launch_r<1, 1024>(a, B, nt, st);
    if (rc)
        return rc;
    ++g_fnftb_launch_count;
    return (int)cudaGetLastError();
}

#endif  // FNFTB_EMUL

#ifndef FNFTB_EMUL
// ---- z -> lambda and bounding-box filter on the device ---------------------------------------
// lambda = log(z) / (i * lam_den)   (fnft__akns_discretization_z_to_lambda,
// src/private/fnft__akns_discretization.c:225-240) followed by misc_filter
// (src/private/fnft__misc.c:114-157): order-preserving compaction of the values inside the box, so
// that only the survivors (tens out of thousands of roots) travel to the host.
struct RootsLamArgs {
    const cplx *roots;   // [B][n]
    const int *info;     // [B][4] (m = info[1] valid roots per polynomial)
    cplx *lam;           // [B][n] compacted
    int *count;          // [B]
    int n;
    double lam_den;
    int filtering;       // 0: keep everything
    double box[4];       // re_min, re_max, im_min, im_max
    const double *box3;  // per-polynomial im_max (overrides box[3]) or NULL
};

__global__ void __launch_bounds__(256) k_roots_lambda(const RootsLamArgs a)
{
    __shared__ int cnt[256];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int m = a.info[4 * b + 1];
    const cplx *z = a.roots + (size_t)b * a.n;
    cplx *out = a.lam + (size_t)b * a.n;
    const double im_max = a.box3 ? a.box3[b] : a.box[3];
    const int per = (m + nt - 1) / nt;
    const int lo = min(tid * per, m), hi = min(lo + per, m);
    const double f = 1.0 / a.lam_den;
    // pass 1: count
    int k = 0;
    for (int i = lo; i < hi; ++i) {
        const cplx zi = z[i];
        const double re = atan2(zi.y, zi.x) * f, im = -log(hypot(zi.x, zi.y)) * f;
        const bool keep = !a.filtering || (re >= a.box[0] && re <= a.box[1] && im >= a.box[2] && im <= im_max);
        k += keep ? 1 : 0;
    }
    cnt[tid] = k;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < nt; ++t) {
            const int c = cnt[t];
            cnt[t] = run;
            run += c;
        }
        a.count[b] = run;
    }
    __syncthreads();
    // pass 2: write in order
    int w = cnt[tid];
    for (int i = lo; i < hi; ++i) {
        const cplx zi = z[i];
        const double re = atan2(zi.y, zi.x) * f, im = -log(hypot(zi.x, zi.y)) * f;
        const bool keep = !a.filtering || (re >= a.box[0] && re <= a.box[1] && im >= a.box[2] && im <= im_max);
        if (keep)
            out[w++] = make_cplx(re, im);
    }
}
#endif

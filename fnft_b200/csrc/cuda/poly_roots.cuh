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
//   * one CTA per polynomial, roots in shared memory (Jacobi sweeps: all updates of a sweep
//     use the previous sweep's roots), coefficients read through the read-only path (all lanes
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
template <int NT, int RMAX>
__global__ void __launch_bounds__(NT) k_roots_aberth(const RootsArgs a)
{
    extern __shared__ double2 fnftb_smem[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = a.n;
    const int lead = a.info[4 * b + 0], m = a.info[4 * b + 1];
    if (m <= 0)
        return;
    // stripped polynomial: coefficients c[lead .. lead+m], highest power first
    const cplx *c = a.coef + (size_t)b * a.cstride + lead;
    const double *ac = a.absc + (size_t)b * (n + 1) + lead;
    cplx *groots = a.roots + (size_t)b * n;
    cplx *z = (cplx *)fnftb_smem;  // [m]
    for (int i = tid; i < m; i += NT)
        z[i] = groots[i];
    __syncthreads();
    unsigned done = 0;  // bit r: root tid + r*NT has converged
    const double tol = 4.0 * (double)m * 2.220446049250313e-16;
    int it = 0, left = m;
    for (; it < a.maxit && left > 0; ++it) {
        cplx znew[RMAX];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const int i = tid + r * NT;
            if (i >= m || ((done >> r) & 1u))
                continue;
            const cplx zi = z[i];
            const double az = hypot(zi.x, zi.y);
            cplx Nw;  // Newton correction p/p'
            bool conv;
            if (az <= 1.0) {
                cplx p = LDG(&c[0]), dp = czero();
                double e = LDG(&ac[0]);
                for (int k = 1; k <= m; ++k) {
                    dp = cadd(cmul(dp, zi), p);
                    p = cadd(cmul(p, zi), LDG(&c[k]));
                    e = e * az + LDG(&ac[k]);
                }
                conv = (hypot(p.x, p.y) <= tol * e);
                Nw = cdiv(p, dp);
            } else {
                const cplx w = cdiv(make_cplx(1.0, 0.0), zi);
                const double aw = 1.0 / az;
                cplx p = LDG(&c[m]), dp = czero();
                double e = LDG(&ac[m]);
                for (int k = m - 1; k >= 0; --k) {
                    dp = cadd(cmul(dp, w), p);
                    p = cadd(cmul(p, w), LDG(&c[k]));
                    e = e * aw + LDG(&ac[k]);
                }
                conv = (hypot(p.x, p.y) <= tol * e);
                // N = z / (m - w q'(w)/q(w))
                const cplx t = cmul(w, cdiv(dp, p));
                Nw = cdiv(zi, make_cplx((double)m - t.x, -t.y));
            }
            if (conv) {  // frozen from now on (it still enters the sums of the others)
                done |= (1u << r);
                continue;
            }
            if (!(isfinite(Nw.x) && isfinite(Nw.y)))  // p' = 0: leave the stationary point sideways
                Nw = make_cplx(1e-3 * az + 1e-6, 1e-3 * az + 1e-6);
            float sx = 0.f, sy = 0.f;
            for (int j = 0; j < m; ++j) {
                const cplx zj = z[j];
                const float dx = (float)(zi.x - zj.x), dy = (float)(zi.y - zj.y);
                const float r2 = fmaxf(dx * dx + dy * dy, 1e-37f);
                const float inv = __frcp_rn(r2);
                sx = fmaf(dx, inv, sx);
                sy = fmaf(-dy, inv, sy);
            }
            const cplx S = make_cplx((double)sx, (double)sy);
            const cplx den = csub(make_cplx(1.0, 0.0), cmul(Nw, S));
            cplx dz = cdiv(Nw, den);
            if (!(isfinite(dz.x) && isfinite(dz.y)))
                dz = Nw;
            znew[r] = csub(zi, dz);
        }
        __syncthreads();
        int mine = 0;
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const int i = tid + r * NT;
            if (i < m) {
                if (!((done >> r) & 1u)) {
                    z[i] = znew[r];
                    ++mine;
                }
            }
        }
        left = __syncthreads_count(mine > 0);
        // __syncthreads_count counts threads, not roots: good enough as a loop condition
    }
    int notconv = 0;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const int i = tid + r * NT;
        if (i < m) {
            if (!((done >> r) & 1u)) {  // still moving after maxit sweeps: not a root, say so
                ++notconv;
                groots[i] = make_cplx(nan(""), nan(""));
            } else {
                groots[i] = z[i];
            }
        }
    }
    if (notconv)
        atomicAdd(&a.info[4 * b + 3], notconv);
    if (tid == 0)
        a.info[4 * b + 2] = it;
}

// Launches both kernels for B polynomials of nominal degree n.  Returns 0, or -6 when n
// exceeds what one CTA can hold (roots in shared memory: n <= 8192).
static inline int roots_launch(const RootsArgs &a, int B, cudaStream_t st)
{
    if (a.n < 1)
        return -2;
    if (a.n > 8192)
        return -6;
    k_roots_init<<<B, 256, 0, st>>>(a);
    ++g_fnftb_launch_count;
    const size_t smem = sizeof(cplx) * (size_t)a.n;
    cudaError_t e = cudaSuccess;
    if (a.n <= 1024) {
        auto kern = k_roots_aberth<256, 4>;
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<B, 256, smem, st>>>(a);
    } else {
        auto kern = k_roots_aberth<1024, 8>;
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return (int)e;
        kern<<<B, 1024, smem, st>>>(a);
    }
    ++g_fnftb_launch_count;
    return (int)cudaGetLastError();
}

#endif  // FNFTB_EMUL

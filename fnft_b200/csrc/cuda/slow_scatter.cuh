// fnft_b200 -- continuous spectrum with the non-polynomial ("slow") discretizations BO, CF4_2, CF4_3, CF5_3, CF6_4,
// ES4 and TES4.
//
// Replaces, for these two discretizations, the call of fnft__nse_scatter_matrix (derivative_flag 0)
//   /root/reference/src/private/fnft__akns_scatter_matrix.c:112-126,206-232
// from nsev_compute_contspec and the epilogue that follows it
//   /root/reference/src/fnft_nsev.c:794-814,836-876.
// O(D*M) work per signal by construction (that is what "slow" means in the reference: one product of
// D step matrices exp([[-i l, q], [r, i l]] eps_t) per spectral point); one warp per (signal, xi):
// the D factors are split into 32 contiguous chunks, multiplied up per lane and combined with an
// ordered shuffle reduction, as in bound_warp.cuh / nsep_refine.cuh.
#pragma once
#ifndef FNFTB_EMUL
#include "nsep_refine.cuh"

struct SlowCsArgs {
    const cplx *q;   // [B][D] effective (preprocessed) samples
    const cplx *r;   // [B][D] explicit r (CF5_3 / CF6_4), NULL: r = -kappa*conj(q)
    int wsel;        // weights of the spectral parameter, see bo_l_at
    int B, D, upsampling, kappa;
    int M, cstype;   // 0: rho, 1: a and b, 2: rho, a, b
    double eps_t, lweight;
    double xi0, eps_xi;
    double ph_rho, ph_a, ph_b;
    cplx *out;       // [B][out_sstride]
    size_t out_sstride;
    int *status;     // [B], 3 = division by zero
};

// product of the steps of samples [lo, hi) without derivative, r = -kappa*conj(q)
DEV void bo_chunk_plain(const cplx *q, const cplx *r, int lo, int hi, cplx l, double h, int kappa, int wsel, cplx *Pm)
{
    Pm[0] = make_cplx(1.0, 0.0);
    Pm[1] = czero();
    Pm[2] = czero();
    Pm[3] = make_cplx(1.0, 0.0);
    const double ks = -(double)kappa;
    for (int n = lo; n < hi;) {
        cplx U[4], Ud[4], t[4];
        n += slow_step_at<false>(q, r, n, l, h, wsel, ks, U, Ud);
        bo_mm(U, Pm, t);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            Pm[i] = t[i];
    }
}

__global__ void __launch_bounds__(128) k_slow_contspec(const SlowCsArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long gid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gid >= (long long)a.B * a.M)
        return;
    const int s = (int)(gid / a.M), m = (int)(gid % a.M);
    const cplx *q = a.q + (size_t)s * a.D;
    int lo, hi;
    bo_chunk_bounds(a.D, a.upsampling, lane, &lo, &hi);
    const double xi = a.xi0 + a.eps_xi * (double)m;
    cplx P[4];
    bo_chunk_plain(q, a.r ? a.r + (size_t)s * a.D : (const cplx *)0, lo, hi, make_cplx(xi * a.lweight, 0.0), a.eps_t,
                   a.kappa, a.wsel, P);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        cplx H[4], t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            H[e] = shfl_down_c(P[e], off);
        if ((lane & (2 * off - 1)) == 0) {
            bo_mm(H, P, t);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                P[e] = t[e];
        }
    }
    if (lane != 0)
        return;
    const cplx H11 = P[0], H21 = P[2];
    cplx *o = a.out + (size_t)s * a.out_sstride;
    size_t off = 0;
    double sn, cs;
    if (a.cstype == 0 || a.cstype == 2) {
        if (H11.x == 0.0 && H11.y == 0.0) {
            a.status[s] = 3;
            o[m] = make_cplx(NAN, NAN);
        } else {
            sincos(xi * a.ph_rho, &sn, &cs);
            o[m] = cdiv(cmul(H21, make_cplx(cs, sn)), H11);
        }
        off = (size_t)a.M;
    }
    if (a.cstype == 1 || a.cstype == 2) {
        sincos(xi * a.ph_a, &sn, &cs);
        o[off + m] = cmul(H11, make_cplx(cs, sn));
        sincos(xi * a.ph_b, &sn, &cs);
        o[off + a.M + m] = cmul(H21, make_cplx(cs, sn));
    }
}
#endif

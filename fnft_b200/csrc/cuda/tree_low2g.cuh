// fnft_b200 -- spectrum-carry kernel for the LOW part of the product tree, GENERAL 2x2 case
// (KdV: r = -1; explicit r of fnft__akns_fscatter).  Same scheme as tree_low2.cuh (see there
// and DESIGN.md 3a) with all four entries of every matrix carried:
//
//  * thread phase: thread t builds the degree-4 matrix of its 4/DEG0 samples in registers and
//    evaluates the four polynomials at the 8th roots of unity (pruned register FFT);
//  * levels N = 8, 16, ..., 4*M: X stage (pending forward pass + full 2x2 pointwise product
//    + first inverse pass; one thread owns ALL eight operand arrays of its 4 positions, so the
//    in-place update needs no barrier), P / M stages exactly as in the symmetric kernel -- the
//    workspace of output array w = 4*pair + entry sits at w*2N + N in both;
//  * output: values of the CTA's degree-4M matrix at the 8M-th roots of unity, or its
//    coefficients when the CTA covers the whole signal.
//
// The top / bottom coefficients travel as plain 2x2 matrices: top(A*B) = top(A)*top(B).
#pragma once
#ifndef FNFTB_EMUL
#include "tree_low2.cuh"

struct GenTops {
    cplx t[4];  // top coefficients (index d) of entries 11, 12, 21, 22
    cplx b[4];  // bottom coefficients (index 0)
};

struct Low2gArgs {
    const cplx *q;   // [B][D]
    const cplx *r;   // [B][D] (rmode explicit) or NULL
    cplx *out;       // spec_out: [B][nblk][4][2N]; else coefficients [B][nblk][4][N+1]
    double *mx_out;  // [B][nblk]
    int *W;          // [B]
    int *status;     // [B]
    GenTops *tt_out; // [B][nblk]
    TwSet tw;
    int spec_out;
    int B, D, npad;
    int rmode, kappa, scheme, normalize;
    double eps_t;
};

DEV GenTops gen_pair_tops(const GenTops &A, const GenTops &B)
{
    GenTops o;
    o.t[0] = cmul(A.t[0], B.t[0]);
    cfma(o.t[0], A.t[1], B.t[2]);
    o.t[1] = cmul(A.t[0], B.t[1]);
    cfma(o.t[1], A.t[1], B.t[3]);
    o.t[2] = cmul(A.t[2], B.t[0]);
    cfma(o.t[2], A.t[3], B.t[2]);
    o.t[3] = cmul(A.t[2], B.t[1]);
    cfma(o.t[3], A.t[3], B.t[3]);
    o.b[0] = cmul(A.b[0], B.b[0]);
    cfma(o.b[0], A.b[1], B.b[2]);
    o.b[1] = cmul(A.b[0], B.b[1]);
    cfma(o.b[1], A.b[1], B.b[3]);
    o.b[2] = cmul(A.b[2], B.b[0]);
    cfma(o.b[2], A.b[3], B.b[2]);
    o.b[3] = cmul(A.b[2], B.b[1]);
    cfma(o.b[3], A.b[3], B.b[3]);
    return o;
}

DEV cplx l2_shfl_down_c(cplx v, int off)
{
    return make_cplx(__shfl_down_sync(0xffffffffu, v.x, off), __shfl_down_sync(0xffffffffu, v.y, off));
}

DEV GenTops gen_tops_scaled(GenTops o, double f)
{
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        o.t[e] = cscale(o.t[e], f);
        o.b[e] = cscale(o.b[e], f);
    }
    return o;
}

// full 2x2 product of matrices of degree D (entries [4][D+1])
template <int D>
DEV void gen_prod(const cplx (*A)[D + 1], const cplx (*B)[D + 1], cplx (*C)[2 * D + 1])
{
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int k = 0; k <= 2 * D; ++k)
            C[e][k] = czero();
#pragma unroll
    for (int row = 0; row < 2; ++row)
#pragma unroll
        for (int col = 0; col < 2; ++col)
#pragma unroll
            for (int i = 0; i <= D; ++i)
#pragma unroll
                for (int j = 0; j <= D; ++j) {
                    cfma(C[row * 2 + col][i + j], A[row * 2 + 0][i], B[0 * 2 + col][j]);
                    cfma(C[row * 2 + col][i + j], A[row * 2 + 1][i], B[1 * 2 + col][j]);
                }
}

template <int DEG0>
__device__ __noinline__ void low2g_leaf_generic(int scheme, double eps_t, double qx, double qy, double rx,
                                                double ry, cplx *P /* [4][DEG0+1] */, int *err)
{
    leaf_matrix(P, scheme, DEG0, eps_t, make_cplx(qx, qy), make_cplx(rx, ry), err);
}

// 2SPLIT4B / 4SPLIT4B leaf (fnft__akns_fscatter.c:402-433) for REAL q*r: Delta = h*sqrt(-q r) is
// real or purely imaginary, so E(h) = (cos D, q h sinc D, r h sinc D) needs one real sincos or
// sinh/cosh of x = (eps/4) sqrt|q r| and the double-angle formulas.  KdV (r = -1, real u) and
// any explicit r with real q*r take this path.
DEV void low2g_leaf_4b_real(double eps_t, cplx q, cplx r, double qr, cplx (*P)[3])
{
    const double ha = 0.25 * eps_t, hb = 0.5 * eps_t;
    const double x = ha * sqrt(fabs(qr));
    double c1, c2, sa, sb;
    if (qr <= 0.0) {  // -q r >= 0: trigonometric
        double sn, cs;
        sincos(x, &sn, &cs);
        c1 = cs;
        c2 = 1.0 - 2.0 * sn * sn;
        const double s2 = 2.0 * sn * cs;
        sa = (x >= 1.0e-8) ? sn / x : 1.0;
        sb = (2.0 * x >= 1.0e-8) ? s2 / (2.0 * x) : 1.0;
    } else {
        const double sh = sinh(x), ch = cosh(x);
        c1 = ch;
        c2 = 1.0 + 2.0 * sh * sh;
        const double s2 = 2.0 * sh * ch;
        sa = (x >= 1.0e-8) ? sh / x : 1.0;
        sb = (2.0 * x >= 1.0e-8) ? s2 / (2.0 * x) : 1.0;
    }
    const double al = ha * sa, be = hb * sb;  // a_1 = q al, a_2 = r al, b_1 = q be, b_2 = r be
    const double third = 1.0 / 3.0;
    const double p0 = (4.0 * c2 * al * al - be * be) * qr * third;
    const double p1 = 8.0 * c1 * al * be * qr * third;
    const double p2 = c2 * (4.0 * c1 * c1 - c2) * third;
    const double g0 = c2 * (4.0 * c1 * al - be) * third;
    // p12[1] = 4 (b_1 a_0^2 + b_2 a_1^2)/3 = (4 be/3) (c1^2 q + al^2 r q^2);  p21[1] symmetric
    const double g1a = 4.0 * be * c1 * c1 * third, g1b = 4.0 * be * al * al * qr * third;
    P[0][0] = make_cplx(p0, 0.0);
    P[0][1] = make_cplx(p1, 0.0);
    P[0][2] = make_cplx(p2, 0.0);
    P[1][0] = cscale(q, g0);
    P[1][1] = cscale(q, g1a + g1b);  // al^2 r q^2 = al^2 (q r) q
    P[1][2] = P[1][0];
    P[2][0] = cscale(r, g0);
    P[2][1] = cscale(r, g1a + g1b);
    P[2][2] = P[2][0];
    P[3][0] = P[0][2];
    P[3][1] = P[0][1];
    P[3][2] = P[0][0];
}

template <int DEG0>
DEV void low2g_leaf(const Low2gArgs &a, int s, int mg, cplx (*P)[DEG0 + 1], int *err)
{
    if (mg < a.D) {
        const size_t idx = (size_t)s * a.D + (size_t)(a.D - 1 - mg);
        const cplx q = a.q[idx];
        cplx r;
        if (a.rmode == FNFTB_R_NSE)
            r = (a.kappa == 1) ? make_cplx(-q.x, q.y) : make_cplx(q.x, -q.y);
        else if (a.rmode == FNFTB_R_KDV)
            r = make_cplx(-1.0, 0.0);
        else
            r = a.r[idx];
        const cplx qr = cmul(q, r);
        if (DEG0 == 2 && qr.y == 0.0 && (a.scheme == FNFTB_AKNS_2SPLIT4B || a.scheme == FNFTB_AKNS_4SPLIT4B)) {
            if constexpr (DEG0 == 2)
                low2g_leaf_4b_real(a.eps_t, q, r, qr.x, P);
        } else {
            cplx tmp[4 * (DEG0 + 1)];
            low2g_leaf_generic<DEG0>(a.scheme, a.eps_t, q.x, q.y, r.x, r.y, tmp, err);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int i = 0; i <= DEG0; ++i)
                    P[e][i] = tmp[e * (DEG0 + 1) + i];
        }
    } else {  // padding z^deg * I (fnft__poly_fmult.c:422-438)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int i = 0; i <= DEG0; ++i)
                P[e][i] = czero();
        P[0][0] = make_cplx(1.0, 0.0);
        P[3][0] = make_cplx(1.0, 0.0);
    }
}

template <int DEG0, int LC>
struct Low2gBuild {
    DEV static void run(const Low2gArgs &a, int s, int mg0, cplx (*P)[(DEG0 << LC) + 1], int *err)
    {
        if constexpr (LC == 0) {
            low2g_leaf<DEG0>(a, s, mg0, P, err);
        } else {
            constexpr int DH = DEG0 << (LC - 1);
            cplx A[4][DH + 1], Bm[4][DH + 1];
            Low2gBuild<DEG0, LC - 1>::run(a, s, mg0, A, err);
            Low2gBuild<DEG0, LC - 1>::run(a, s, mg0 + (1 << (LC - 1)), Bm, err);
            gen_prod<DH>(A, Bm, P);
        }
    }
};

// values of x (degree 4, 5 coefficients) at the 8th roots of unity -> S[base + brev3(k)]
DEV void low2g_front_fft8(const cplx *x, cplx *S, int base)
{
    cplx e[4], o[4];
    e[0] = x[0];
    o[0] = x[0];
    e[1] = x[1];
    o[1] = mul_root<8, 1, -1>(x[1]);
    e[2] = x[2];
    o[2] = mul_root<8, 2, -1>(x[2]);
    e[3] = x[3];
    o[3] = mul_root<8, 3, -1>(x[3]);
    Dft<4, -1>::run(e);
    Dft<4, -1>::run(o);
    const cplx t = x[4];
    const int ad = swz2(base);  // base is a multiple of 8
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        S[ad ^ brev_c(m, 2)] = cadd(e[m], t);
        S[ad ^ (4 + brev_c(m, 2))] = csub(o[m], t);
    }
}

// X stage, general 2x2: thread t owns (pair p, positions 4g .. 4g+3) of all eight operand
// arrays; output array e of the pair goes to p*8N + e*2N (even bins) and + N (workspace).
template <int LOG2M>
DEV void low2g_x_stage(cplx *S, const GenTops *TTc, int t, int l2n, bool first, bool last, cplx *gout)
{
    const int N = 1 << l2n;
    const int l2g = l2n - 2;  // groups per array
    const int p = t >> l2g, g = t & ((1 << l2g) - 1);
    const bool odd = (4 * g >= (N >> 1));
    const int base = (p << (l2n + 3)) + 4 * g;
    int ad[8];
#pragma unroll
    for (int arr = 0; arr < 8; ++arr)
        ad[arr] = swz2(base + (arr << l2n));
    cplx v[8][4];
#pragma unroll
    for (int arr = 0; arr < 8; ++arr)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            v[arr][j] = S[ad[arr] ^ j];
    if (odd && !first) {
#pragma unroll
        for (int arr = 0; arr < 8; ++arr) {
            Dft<4, -1>::run(v[arr]);
            const cplx t1 = v[arr][1];
            v[arr][1] = v[arr][2];
            v[arr][2] = t1;
            const cplx ct = TTc[2 * p + (arr >> 2)].t[arr & 3];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[arr][j] = csub(v[arr][j], ct);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int row = e >> 1, col = e & 1;
        cplx c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c[j] = cmul(v[row * 2][j], v[4 + col][j]);
            cfma(c[j], v[row * 2 + 1][j], v[4 + 2 + col][j]);
        }
        // output array e: even bins at offset e*2N, workspace at e*2N + N  (the same eight
        // slots per position this thread has just read)
        if (!last) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                S[ad[2 * e] ^ j] = c[j];
        } else if (gout) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                gout[((size_t)e << (l2n + 1)) + 4 * g + j] = c[j];
        }
        const cplx t1 = c[1];
        c[1] = c[2];
        c[2] = t1;
        Dft<4, +1>::run(c);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            S[ad[2 * e + 1] ^ j] = c[j];
    }
}

// M stage on workspace array wv (at wv*2N + N); bots[wv] is c_0 of that output polynomial
template <int LOG2M, int R>
DEV double low2g_m_stage(cplx *S, const TwSet &tw, const GenTops *TTn, int t, int l2n, int sb, bool want_max)
{
    constexpr int M = 1 << LOG2M;
    constexpr int LR = Log2R<R>::value;
    constexpr int PER = 16 / R;
    const int l2s = l2n - LR;
    const int s = 1 << l2s;
    const double invN = 1.0 / (double)(1 << l2n);
    const cplx *pt = tw.base + tw.pass_off[l2n][LR];
    const cplx *tt = tw.base + tw.twist_off[l2n];
    double m2 = 0.0;
#pragma unroll 1
    for (int k = 0; k < PER; ++k) {
        int idx = t + k * M;
        if (l2s == 2)
            idx = swapbit2(idx, sb);
        const int o = idx & (s - 1);
        const int wv = idx >> l2s;
        const int base = (wv << (l2n + 1)) + (1 << l2n) + o;
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q)
            v[q] = S[swz2(base + (brev_c(q, LR) << l2s))];
        up_twiddle_mul<R, true>(v, pt, s, o);
        Dft<R, +1>::run(v);
#pragma unroll
        for (int n = 0; n < R; ++n)
            v[n] = cscale(v[n], invN);
        if (o == 0)
            v[0] = TTn[wv >> 2].b[wv & 3];
        if (want_max) {
#pragma unroll
            for (int n = 0; n < R; ++n)
                m2 = fmax(m2, cabs2(v[n]));
        }
        UpTwist<R, 0>::run(v, __ldg(&tt[o]));
        Dft<R, -1>::run(v);
        up_twiddle_mul<R, false>(v, pt, s, o);
#pragma unroll
        for (int q = 0; q < R; ++q)
            S[swz2(base + (brev_c(q, LR) << l2s))] = v[q];
    }
    return m2;
}

// coefficient output of the single remaining matrix: last inverse pass (radix 16), 1/N
template <int LOG2M>
DEV double low2g_out_stage(const cplx *S, const TwSet &tw, const GenTops &Tn, int t, int l2n, cplx *out)
{
    constexpr int M = 1 << LOG2M;
    constexpr int R = 16, LR = 4;
    const int l2s = l2n - LR;
    const int s = 1 << l2s;
    const int N = 1 << l2n;
    const double invN = 1.0 / (double)N;
    const cplx *pt = tw.base + tw.pass_off[l2n][LR];
    double m2 = 0.0;
#pragma unroll 1
    for (int idx = t; idx < 4 * s; idx += M) {
        const int o = idx & (s - 1);
        const int e = idx >> l2s;
        const int base = (e << (l2n + 1)) + N + o;
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q)
            v[q] = S[swz2(base + (brev_c(q, LR) << l2s))];
        up_twiddle_mul<R, true>(v, pt, s, o);
        Dft<R, +1>::run(v);
        cplx *dst = out + (size_t)e * (N + 1);
#pragma unroll
        for (int n = 0; n < R; ++n) {
            cplx c = cscale(v[n], invN);
            if (n == 0 && o == 0)
                c = Tn.b[e];
            dst[o + (n << l2s)] = c;
            m2 = fmax(m2, cabs2(c));
        }
        if (o == 0) {
            dst[N] = Tn.t[e];
            m2 = fmax(m2, cabs2(Tn.t[e]));
        }
    }
    return m2;
}

// spec_out: last forward pass (radix 4, stride 1) of the four workspace arrays, "- c_N"
template <int LOG2M>
DEV void low2g_final_stage(const cplx *S, const GenTops &Tn, int t, int l2n, cplx *gout)
{
    constexpr int M = 1 << LOG2M;
    const int N = 1 << l2n;
#pragma unroll 1
    for (int it = t; it < N; it += M) {  // 4 arrays * N/4 groups
        const int e = it >> (l2n - 2);
        const int g = it & ((N >> 2) - 1);
        const int ad = swz2((e << (l2n + 1)) + N + 4 * g);
        cplx v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            v[j] = S[ad ^ j];
        Dft<4, -1>::run(v);
        const cplx ct = Tn.t[e];
        cplx *dst = gout + ((size_t)e << (l2n + 1)) + N + 4 * g;
        dst[0] = csub(v[0], ct);
        dst[1] = csub(v[2], ct);
        dst[2] = csub(v[1], ct);
        dst[3] = csub(v[3], ct);
    }
}

DEV void low2g_plan(int l2n, int *rp, int *rm, int *sbp, int *sbm)
{
    if (l2n == 3) {
        *rp = 0;
        *rm = 2;
        *sbp = 2;
        *sbm = 3;
        return;
    }
    low2_plan(l2n, rp, rm, sbp, sbm);
}

// grid.x = B * (npad / S), S = M * 4 / DEG0, blockDim.x = M
template <int LOG2M, int DEG0>
__global__ void __launch_bounds__(1 << LOG2M, 1) k_tree_low2g(const Low2gArgs a)
{
    constexpr int M = 1 << LOG2M;
    constexpr int LC = (DEG0 == 2) ? 1 : 2;  // log2(samples per thread)
    constexpr int S_ = M << LC;
    extern __shared__ double2 fnftb_smem2g[];
    cplx *S = (cplx *)fnftb_smem2g;             // 32*M
    GenTops *TT0 = (GenTops *)(S + 32 * M);     // M/2
    GenTops *TT1 = TT0 + M / 2;                 // M/4
    double *red = (double *)(TT1 + M / 4);
    const int t = threadIdx.x;
    const int nblk = a.npad / S_;
    const int s = blockIdx.x / nblk, blk = blockIdx.x % nblk;

    // ---- thread phase ------------------------------------------------------------------
    GenTops mine;
    {
        cplx P[4][5];
        int err = 0;
        Low2gBuild<DEG0, LC>::run(a, s, blk * S_ + (t << LC), P, &err);
        if (err && a.status)
            a.status[s] = err;
        int ex = 0;
        if (a.normalize) {
            double m = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    m = fmax(m, fmax(fabs(P[e][i].x), fabs(P[e][i].y)));
            ex = rescale_exponent(m);
            if (ex != 0) {
                const double sc = ldexp(1.0, -ex);
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int i = 0; i < 5; ++i)
                        P[e][i] = cscale(P[e][i], sc);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
                ex += __shfl_xor_sync(0xffffffffu, ex, off);
            if ((t & 31) == 0 && ex != 0)
                atomicAdd(&a.W[s], ex);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            mine.t[e] = P[e][4];
            mine.b[e] = P[e][0];
            low2g_front_fft8(P[e], S, t * 32 + e * 8);
        }
    }
    {
        GenTops nb;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            nb.t[e] = l2_shfl_down_c(mine.t[e], 1);
            nb.b[e] = l2_shfl_down_c(mine.b[e], 1);
        }
        if ((t & 1) == 0)
            TT0[t >> 1] = gen_pair_tops(mine, nb);
    }
    __syncthreads();

    // ---- levels: operand length 8, 16, ..., 4*M ------------------------------------------
    GenTops *TTc = TT1, *TTn = TT0;
    const size_t mo = (size_t)s * nblk + blk;
    cplx *gspec = a.spec_out ? a.out + mo * (size_t)(32 * M) : nullptr;  // 4 arrays of 8*M
    double m2 = 0.0;
#pragma unroll 1
    for (int L = 0; L < LOG2M; ++L) {
        const int l2n = 3 + L;
        const bool last = (L == LOG2M - 1);
        int rp, rm, sbp, sbm;
        low2g_plan(l2n, &rp, &rm, &sbp, &sbm);
        if (L > 0 && t < (M >> (L + 1)))
            TTn[t] = gen_pair_tops(TTc[2 * t], TTc[2 * t + 1]);
        low2g_x_stage<LOG2M>(S, TTc, t, l2n, L == 0, last, gspec);
        __syncthreads();
        switch (rp) {
        case 4: low2_p_pass<LOG2M, 4, +1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 8: low2_p_pass<LOG2M, 8, +1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 16: low2_p_pass<LOG2M, 16, +1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        default: break;
        }
        if (last && !a.spec_out) {
            m2 = low2g_out_stage<LOG2M>(S, a.tw, TTn[0], t, l2n, a.out + mo * 4 * ((4 * M) + 1));
            break;
        }
        const bool want_max = last;
        double mm;
        switch (rm) {
        case 2: mm = low2g_m_stage<LOG2M, 2>(S, a.tw, TTn, t, l2n, sbm, want_max); break;
        case 4: mm = low2g_m_stage<LOG2M, 4>(S, a.tw, TTn, t, l2n, sbm, want_max); break;
        case 8: mm = low2g_m_stage<LOG2M, 8>(S, a.tw, TTn, t, l2n, sbm, want_max); break;
        default: mm = low2g_m_stage<LOG2M, 16>(S, a.tw, TTn, t, l2n, sbm, want_max); break;
        }
        m2 = fmax(m2, mm);
        __syncthreads();
        switch (rp) {
        case 4: low2_p_pass<LOG2M, 4, -1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 8: low2_p_pass<LOG2M, 8, -1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 16: low2_p_pass<LOG2M, 16, -1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        default: break;
        }
        if (last) {
            low2g_final_stage<LOG2M>(S, TTn[0], t, l2n, gspec);
            if (t == 0) {
                a.tt_out[mo] = TTn[0];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    m2 = fmax(m2, cabs2(TTn[0].t[e]));
            }
            break;
        }
        GenTops *tmp = TTc;
        TTc = TTn;
        TTn = tmp;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, off));
    if ((t & 31) == 0)
        red[t >> 5] = m2;
    __syncthreads();
    if (t == 0) {
        double m = red[0];
        for (int w = 1; w < M / 32; ++w)
            m = fmax(m, red[w]);
        a.mx_out[mo] = sqrt(m);
    }
}

static inline size_t low2g_smem_bytes(int log2m)
{
    const size_t M = (size_t)1 << log2m;
    return sizeof(cplx) * 32 * M + sizeof(GenTops) * (M / 2 + M / 4) + sizeof(double) * 16;
}
static inline int low2g_samples(int log2m, int deg0) { return (1 << log2m) * (4 / deg0); }

#ifdef FNFTB_TU_LOW2
template <int LOG2M, int DEG0>
static inline int low2g_launch_t(const Low2gArgs &a, cudaStream_t st)
{
    const size_t smem = low2g_smem_bytes(LOG2M);
    const int e = fnftb_smem_optin((const void *)k_tree_low2g<LOG2M, DEG0>, smem);
    if (e != 0)
        return e;
    const unsigned grid = (unsigned)a.B * (unsigned)(a.npad / low2g_samples(LOG2M, DEG0));
    if (g_fnftb_profile_on)
        fnftb_profile_begin("tree_low2g", st);
    k_tree_low2g<LOG2M, DEG0><<<grid, 1 << LOG2M, smem, st>>>(a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(st);
    ++g_fnftb_launch_count;
    return (int)cudaGetLastError();
}
// the one shape the driver uses: 256 threads per CTA
int low2g_launch(const Low2gArgs &a, int deg0, cudaStream_t st)
{
    return (deg0 == 2) ? low2g_launch_t<8, 2>(a, st) : low2g_launch_t<8, 1>(a, st);
}
#else
int low2g_launch(const Low2gArgs &a, int deg0, cudaStream_t st);
#endif
#endif  // !FNFTB_EMUL

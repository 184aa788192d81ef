// fnft_b200 -- "spectrum carry" kernel for the LOW part of the product tree in the
// first-row-only (NSE) mode: one CTA turns S consecutive samples of one signal into their
// transfer matrix (degree DEG0*S) without leaving shared memory.
//
// Same mathematics as tree_kernels.cuh (leaf: fnft__akns_fscatter.c:116-433; pair product:
// fnft__poly_fmult.c:239-328); what differs from tree_low_kernel.cuh is HOW the FFT
// products are organised:
//
//  * Thread phase.  Thread t builds the degree-8 matrix (a, b) of its 8/DEG0 samples in
//    registers (leaves + direct products), normalises it by a power of two and evaluates it
//    at the 16th roots of unity with a register-resident, zero-pruned 16-point transform.
//  * Every later level keeps the matrices as VALUES at the N-th roots of unity (N = twice
//    the degree), in bit-reversed order.  A pair product is then a pointwise 2x2 product;
//    its N values are exactly the EVEN bins of the length-2N spectrum the next level needs
//    (no transform, no wrap correction).  The ODD bins are FFT_N(c_i * w_2N^i) - c_N, where
//    the coefficients c come from IFFT_N of the product values.  Per polynomial and level
//    this is one inverse and one forward transform of length N instead of one inverse of
//    length N plus one zero-padded forward of length 2N (2/3 of the arithmetic, half of
//    the shared-memory traffic), and coefficients only ever exist in registers.
//  * Stage fusion.  X stage = last forward pass (stride 1) + pointwise product + first
//    inverse pass (stride 1); M stage = last inverse pass + twist by w_2N^i/N + first forward
//    pass, both on the same register set.  In between at most one radix-4/8/16 pass (P).
//  * No rescaling inside the kernel: from matrices normalised to max|c| < 4 the 6 (7)
//    levels cannot overflow (bound 2(d+1)*max^2 per level), and powers of two commute with
//    everything.  The exact max|c| of the CTA's result goes to mx_out for the lazy
//    normalisation of the upper levels (tree_kernels.cuh header).
//
// Shared memory: 32*M cplx spectra (M = matrices per CTA = threads) with the XOR swizzle
// swz2, which together with the lane permutations chosen per stage makes every 16-byte
// access conflict free (checked exhaustively by scripts/low2_banks.py), plus the top /
// bottom coefficients of the current matrices.
#pragma once
#ifndef FNFTB_EMUL
#include "tree_kernels.cuh"
#include "tw_tables.cuh"

struct Low2Tops {
    cplx ta, tb, ba, bb;  // top (index d) and bottom (index 0) coefficients of a and b
};

struct Low2Args {
    const cplx *q;   // [B][D]
    cplx *out;       // [B][npad/S] matrices of degree DEG0*S, entries (a, b)
    double *mx_out;  // [B][npad/S]
    int *W;          // [B]
    int *status;     // [B]
    TwSet tw;        // pass-major twiddle tables (tw_tables.cuh)
    Low2Tops *tt_out;  // [B][npad/S] tops of the results (spec_out only)
    int spec_out;    // 1: write the result as values at the 2*deg roots of unity (bit-reversed),
                     //    layout [a: 2N][b: 2N] per matrix, for tree_up.cuh; 0: coefficients
    int B, D, npad;
    int kappa, scheme, normalize;
    double eps_t;
};

DEV int swz2(int i) { return i ^ (((i >> 3) ^ (i >> 6)) & 7); }

HD constexpr int brev_c(int x, int bits)
{
    int r = 0;
    for (int i = 0; i < bits; ++i)
        r = (r << 1) | ((x >> i) & 1);
    return r;
}

DEV int swapbit2(int x, int sb)
{
    const int b = ((x >> 2) ^ (x >> sb)) & 1;
    return x ^ ((b << 2) | (b << sb));
}

// acc += x * conj(y)
DEV void cfmac(cplx &acc, cplx x, cplx y)
{
    acc.x += x.x * y.x + x.y * y.y;
    acc.y += x.y * y.x - x.x * y.y;
}

// (ao, bo) = first row of [a1 b1; -k b1# a1#] * [a2 b2; -k b2# a2#], degree D each
template <int D>
DEV void sym_prod(const cplx *a1, const cplx *b1, const cplx *a2, const cplx *b2, double kap,
                  cplx *ao, cplx *bo)
{
    cplx b1k[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i)
        b1k[i] = cscale(b1[i], -kap);
#pragma unroll
    for (int k = 0; k <= 2 * D; ++k) {
        ao[k] = czero();
        bo[k] = czero();
    }
#pragma unroll
    for (int i = 0; i <= D; ++i) {
#pragma unroll
        for (int j = 0; j <= D; ++j) {
            cfma(ao[i + j], a1[i], a2[j]);
            cfmac(ao[i + j], b1k[i], b2[D - j]);
            cfma(bo[i + j], a1[i], b2[j]);
            cfmac(bo[i + j], b1[i], a2[D - j]);
        }
    }
}

// v[q] *= w^q (CONJ: conj(w)^q), q = 1 .. R-1, where w = tw[o] is the first row of the pass
// table: the powers are derived in registers (squarings / products of depth <= 4, a few ulp)
// instead of loading R-1 table rows.  ptxas serialises load -> use -> load under register
// pressure, so every one of the 15 L2 latencies of a radix-16 butterfly was exposed
// (long_scoreboard 40-58 % of the stalls, profiles/r01h_kernels_ncu.md); FP64 had headroom.
#ifndef FNFTB_TW_DERIVE
#define FNFTB_TW_DERIVE 1
#endif
DEV cplx csq(cplx a) { return make_cplx(a.x * a.x - a.y * a.y, 2.0 * a.x * a.y); }
template <int R, bool CONJ>
DEV void up_twiddle_mul(cplx *v, const cplx *pt, int stride, int o)
{
#if FNFTB_TW_DERIVE
    if constexpr (R > 16) {
        // radix 32 / 64 (top levels of the longest products only): plain table rows [q-1][o]
#pragma unroll
        for (int q0 = 1; q0 < R; q0 += 8) {
            cplx w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (q0 + j < R)
                    w[j] = __ldg(&pt[(size_t)(q0 + j - 1) * stride + o]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (q0 + j < R)
                    v[q0 + j] = CONJ ? cmulc(v[q0 + j], w[j]) : cmul(v[q0 + j], w[j]);
        }
        return;
    }
    cplx w1 = __ldg(&pt[o]);
    if (CONJ)
        w1 = cconj(w1);
    v[1] = cmul(v[1], w1);
    if constexpr (R > 2) {
        const cplx w2 = csq(w1);
        v[2] = cmul(v[2], w2);
        const cplx w3 = cmul(w2, w1);
        v[3] = cmul(v[3], w3);
        if constexpr (R > 4) {
            const cplx w4 = csq(w2);
            v[4] = cmul(v[4], w4);
            const cplx w5 = cmul(w4, w1);
            v[5] = cmul(v[5], w5);
            const cplx w6 = csq(w3);
            v[6] = cmul(v[6], w6);
            const cplx w7 = cmul(w4, w3);
            v[7] = cmul(v[7], w7);
            if constexpr (R > 8) {
                const cplx w8 = csq(w4);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], csq(w5));
                v[11] = cmul(v[11], cmul(w8, w3));
                v[12] = cmul(v[12], csq(w6));
                v[13] = cmul(v[13], cmul(w8, w5));
                v[14] = cmul(v[14], csq(w7));
                v[15] = cmul(v[15], cmul(w8, w7));
                static_assert(R <= 64, "radix");
            }
        }
    }
    (void)stride;
#else
    constexpr int BATCH = (R > 8) ? 8 : R - 1;
#pragma unroll
    for (int q0 = 1; q0 < R; q0 += BATCH) {
        cplx w[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
            if (q0 + j < R)
                w[j] = __ldg(&pt[(size_t)(q0 + j - 1) * stride + o]);
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
            if (q0 + j < R)
                v[q0 + j] = CONJ ? cmulc(v[q0 + j], w[j]) : cmul(v[q0 + j], w[j]);
    }
#endif
}

// v[n] *= t0 * w_2R^n: the twist w_2N^(o + n*N/R) with t0 = w_2N^o loaded once
template <int R, int N_>
struct UpTwist {
    DEV static void run(cplx *v, cplx t0)
    {
        if constexpr (N_ < R) {
            v[N_] = cmul(mul_root<2 * R, N_, -1>(v[N_]), t0);
            UpTwist<R, N_ + 1>::run(v, t0);
        }
    }
};
DEV void up_twist16(cplx *v, cplx t0) { UpTwist<16, 0>::run(v, t0); }

// generic leaf (any scheme the leaf switch knows): kept out of line, it is large
template <int DEG0>
__device__ __noinline__ void low2_leaf_generic(int scheme, double eps_t, double qx, double qy, double kap,
                                               cplx *AB /* [2][DEG0+1] */, int *err)
{
    const cplx q = make_cplx(qx, qy);
    const cplx r = (kap > 0.0) ? make_cplx(-qx, qy) : make_cplx(qx, -qy);
    cplx p[4 * (DEG0 + 1)];
    leaf_matrix(p, scheme, DEG0, eps_t, q, r, err);
#pragma unroll
    for (int i = 0; i < 2 * (DEG0 + 1); ++i)
        AB[i] = p[i];
}

// 2SPLIT4B / 4SPLIT4B leaf for r = -kappa*conj(q) (fnft__akns_fscatter.c:402-433 with
// -q*r = kappa*|q|^2 real): Delta = h*|q| (kappa = +1) or i*h*|q| (kappa = -1), so the
// zero-frequency exponentials E(h) = (cos D, q*h*sinc D, r*h*sinc D) need one real
// sincos / sinh-cosh of x = eps*|q|/4 and the double-angle formulas for 2x.  With
// a_1 a_2 = -kappa*alpha^2*|q|^2 the entry p11 has real coefficients and p12 = q * real.
DEV void low2_leaf_4b(double eps_t, cplx q, double kap, cplx *A, cplx *B)
{
    const double m2 = q.x * q.x + q.y * q.y;
    const double m = sqrt(m2);
    const double ha = 0.25 * eps_t, hb = 0.5 * eps_t;
    const double x = ha * m;
    double c1, c2, sa, sb;  // cos D_a, cos D_b, sinc D_a, sinc D_b
    if (kap > 0.0) {
        double sn, cs;
        sincos(x, &sn, &cs);
        c1 = cs;
        c2 = 1.0 - 2.0 * sn * sn;
        const double s2 = 2.0 * sn * cs;
        sa = (x >= 1.0e-8) ? sn / x : 1.0;            // misc_CSINC threshold (fnft__misc.c:306-314)
        sb = (2.0 * x >= 1.0e-8) ? s2 / (2.0 * x) : 1.0;
    } else {
        const double sh = sinh(x), ch = cosh(x);
        c1 = ch;
        c2 = 1.0 + 2.0 * sh * sh;
        const double s2 = 2.0 * sh * ch;
        sa = (x >= 1.0e-8) ? sh / x : 1.0;
        sb = (2.0 * x >= 1.0e-8) ? s2 / (2.0 * x) : 1.0;
    }
    const double al = ha * sa, be = hb * sb;           // a_1 = q*al, a_2 = r*al, b_1 = q*be, b_2 = r*be
    const double qr = -kap * m2;                       // q*r
    const double third = 1.0 / 3.0;
    A[0] = make_cplx((4.0 * c2 * al * al - be * be) * qr * third, 0.0);
    A[1] = make_cplx(8.0 * c1 * al * be * qr * third, 0.0);
    A[2] = make_cplx(c2 * (4.0 * c1 * c1 - c2) * third, 0.0);
    const double g0 = c2 * (4.0 * c1 * al - be) * third;
    const double g1 = 4.0 * be * (c1 * c1 + al * al * qr) * third;
    B[0] = make_cplx(q.x * g0, q.y * g0);
    B[1] = make_cplx(q.x * g1, q.y * g1);
    B[2] = B[0];
}

template <int DEG0>
DEV void low2_leaf(const Low2Args &a, int s, int mg, cplx *A, cplx *B, int *err)
{
    if (mg < a.D) {
        const cplx q = a.q[(size_t)s * a.D + (size_t)(a.D - 1 - mg)];
        if (DEG0 == 2 && (a.scheme == FNFTB_AKNS_2SPLIT4B || a.scheme == FNFTB_AKNS_4SPLIT4B)) {
            low2_leaf_4b(a.eps_t, q, (double)a.kappa, A, B);
        } else {
            cplx AB[2 * (DEG0 + 1)];
            low2_leaf_generic<DEG0>(a.scheme, a.eps_t, q.x, q.y, (double)a.kappa, AB, err);
#pragma unroll
            for (int i = 0; i <= DEG0; ++i) {
                A[i] = AB[i];
                B[i] = AB[DEG0 + 1 + i];
            }
        }
    } else {  // padding diag(z^deg, 1) keeps the structure (tree_kernels.cuh header)
#pragma unroll
        for (int i = 0; i <= DEG0; ++i) {
            A[i] = czero();
            B[i] = czero();
        }
        A[0] = make_cplx(1.0, 0.0);
    }
}

// product of 2^LC consecutive level-0 matrices starting at mg0 (degree DEG0 << LC)
template <int DEG0, int LC>
struct Low2Build {
    DEV static void run(const Low2Args &a, int s, int mg0, cplx *A, cplx *B, int *err)
    {
        if constexpr (LC == 0) {
            low2_leaf<DEG0>(a, s, mg0, A, B, err);
        } else {
            constexpr int DH = DEG0 << (LC - 1);
            cplx A1[DH + 1], B1[DH + 1], A2[DH + 1], B2[DH + 1];
            Low2Build<DEG0, LC - 1>::run(a, s, mg0, A1, B1, err);
            Low2Build<DEG0, LC - 1>::run(a, s, mg0 + (1 << (LC - 1)), A2, B2, err);
            sym_prod<DH>(A1, B1, A2, B2, (double)a.kappa, A, B);
        }
    }
};

// values of x (degree 8, 9 coefficients) at the 16th roots of unity -> S[base + brev4(k)]
DEV void low2_front_fft16(const cplx *x, cplx *S, int base)
{
    cplx e[8], o[8];
    e[0] = x[0];
    o[0] = x[0];
    e[1] = x[1];
    o[1] = mul_root<16, 1, -1>(x[1]);
    e[2] = x[2];
    o[2] = mul_root<16, 2, -1>(x[2]);
    e[3] = x[3];
    o[3] = mul_root<16, 3, -1>(x[3]);
    e[4] = x[4];
    o[4] = mul_root<16, 4, -1>(x[4]);
    e[5] = x[5];
    o[5] = mul_root<16, 5, -1>(x[5]);
    e[6] = x[6];
    o[6] = mul_root<16, 6, -1>(x[6]);
    e[7] = x[7];
    o[7] = mul_root<16, 7, -1>(x[7]);
    Dft<8, -1>::run(e);
    Dft<8, -1>::run(o);
    const cplx t = x[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        S[swz2(base + brev_c(m, 3))] = cadd(e[m], t);
        S[swz2(base + 8 + brev_c(m, 3))] = csub(o[m], t);
    }
}

// tops of the pair product A*B (fnft__poly_fmult.c pair product restricted to the extreme
// coefficients), times the pending factor f
DEV Low2Tops low2_pair_tops(const Low2Tops &A, const Low2Tops &B, double kap, double f)
{
    Low2Tops r;
    r.ta = cmul(A.ta, B.ta);
    cfmac(r.ta, cscale(A.tb, -kap), B.bb);
    r.tb = cmul(A.ta, B.tb);
    cfmac(r.tb, A.tb, B.ba);
    r.ba = cmul(A.ba, B.ba);
    cfmac(r.ba, cscale(A.bb, -kap), B.tb);
    r.bb = cmul(A.ba, B.bb);
    cfmac(r.bb, A.bb, B.ta);
    r.ta = cscale(r.ta, f);
    r.tb = cscale(r.tb, f);
    r.ba = cscale(r.ba, f);
    r.bb = cscale(r.bb, f);
    return r;
}

// ---------------------------------------------------------------------------------------
// X stage at operand length N = 1 << l2n.  Item = (pair p, 4 consecutive positions); M
// items in the even half-region (h = 0) and M in the odd one (h = 1), thread t takes item
// t of each.  FIRST: operands are complete (no pending forward pass); LAST: the product
// values are only needed as the input of the inverse transform.
// ---------------------------------------------------------------------------------------
template <int LOG2M>
DEV void low2_x_stage(cplx *S, const Low2Tops *TTc, int t, int l2n, bool first, bool last, double kap,
                          cplx *gout)
{
    const int N = 1 << l2n;
    const int l2g = l2n - 3;
    const int p = t >> l2g, gp = t & ((1 << l2g) - 1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int g = (h << l2g) + gp;
        const int base = (p << (l2n + 2)) + 4 * g;
        int ad[4];
#pragma unroll
        for (int arr = 0; arr < 4; ++arr)
            ad[arr] = swz2(base + (arr << l2n));
        cplx v[4][4];
#pragma unroll
        for (int arr = 0; arr < 4; ++arr)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[arr][j] = S[ad[arr] ^ j];
        if (h == 1 && !first) {
            // pending last forward pass (radix 4, stride 1) and "- c_N" of the odd bins
#pragma unroll
            for (int arr = 0; arr < 4; ++arr) {
                Dft<4, -1>::run(v[arr]);
                const cplx t1 = v[arr][1];
                v[arr][1] = v[arr][2];
                v[arr][2] = t1;
                const Low2Tops &T = TTc[2 * p + (arr >> 1)];
                const cplx ct = (arr & 1) ? T.tb : T.ta;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[arr][j] = csub(v[arr][j], ct);
            }
        }
        // pointwise product; B21 = -kappa*(-1)^k conj(B12), B22 = (-1)^k conj(B11)
        const double sg = h ? -1.0 : 1.0;
        cplx ca[4], cb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const cplx bAk = cscale(v[1][j], -kap * sg);
            const cplx bAs = cscale(v[1][j], sg);
            ca[j] = cmul(v[0][j], v[2][j]);
            cfmac(ca[j], bAk, v[3][j]);
            cb[j] = cmul(v[0][j], v[3][j]);
            cfmac(cb[j], bAs, v[2][j]);
        }
        if (!last) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                S[ad[0] ^ j] = ca[j];
                S[ad[2] ^ j] = cb[j];
            }
        } else if (gout) {  // even bins of the result spectrum (single pair: p == 0)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                gout[4 * g + j] = ca[j];
                gout[(2 << l2n) + 4 * g + j] = cb[j];
            }
        }
        // first inverse pass (radix 4, stride 1, bit-reversed input)
        cplx t1 = ca[1];
        ca[1] = ca[2];
        ca[2] = t1;
        t1 = cb[1];
        cb[1] = cb[2];
        cb[2] = t1;
        Dft<4, +1>::run(ca);
        Dft<4, +1>::run(cb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            S[ad[1] ^ j] = ca[j];
            S[ad[3] ^ j] = cb[j];
        }
    }
}

// workspace array `which` (0: a', 1: b') of pair p at operand length N
DEV int low2_wbase(int p, int which, int l2n)
{
    return (p << (l2n + 2)) + (1 << l2n) + (which << (l2n + 1));
}

// ---------------------------------------------------------------------------------------
// P pass: radix R at stride 4 on the workspace arrays.  DIR = +1: inverse (DIT, bit-reversed
// input), DIR = -1: forward (DIF, bit-reversed output).
// ---------------------------------------------------------------------------------------
// FNFTB_LOW2_BATCH (experiment, off): a thread that owns several butterflies of a stage (radix 4: four, radix 8: two)
// loads all of them, transforms all of them and stores all of them, instead of one after the other.  Same arithmetic,
// same results, same 168 registers -- and SLOWER on B200: 15.50 ms per 4096 signals with the serial loops, 16.79 with
// batched P passes (= 1), 18.83 with batched P passes and M stages (= 2): the serial loop body is a quarter of the
// code and the three resident CTAs already give the scheduler independent chains (round 2, gpurun_out/r5c).
#ifndef FNFTB_LOW2_BATCH
#define FNFTB_LOW2_BATCH 0
#endif
template <int LOG2M, int R, int DIR>
DEV void low2_p_pass(cplx *S, const TwSet &tw, int t, int l2n, int sb)
{
    constexpr int M = 1 << LOG2M;
    constexpr int LR = Log2R<R>::value;
    constexpr int PER = 16 / R;  // items per thread
    const int l2gr = l2n - 2 - LR;  // groups per array (log2)
    const cplx *pt = tw.base + tw.pass_off[LR + 2][LR];  // [q-1][o], o < 4
#if FNFTB_LOW2_BATCH
    cplx v[PER][R];
    int bases[PER], os[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int idx = swapbit2(t + k * M, sb);
        os[k] = idx & 3;
        const int rest = idx >> 2;
        const int g = rest & ((1 << l2gr) - 1);
        const int wv = rest >> l2gr;
        bases[k] = low2_wbase(wv >> 1, wv & 1, l2n) + (g << (LR + 2)) + os[k];
        if (DIR > 0) {
#pragma unroll
            for (int q = 0; q < R; ++q)
                v[k][q] = S[swz2(bases[k] + 4 * brev_c(q, LR))];
        } else {
#pragma unroll
            for (int n = 0; n < R; ++n)
                v[k][n] = S[swz2(bases[k] + 4 * n)];
        }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if (DIR > 0) {
            up_twiddle_mul<R, true>(v[k], pt, 4, os[k]);
            Dft<R, +1>::run(v[k]);
        } else {
            Dft<R, -1>::run(v[k]);
            up_twiddle_mul<R, false>(v[k], pt, 4, os[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if (DIR > 0) {
#pragma unroll
            for (int n = 0; n < R; ++n)
                S[swz2(bases[k] + 4 * n)] = v[k][n];
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q)
                S[swz2(bases[k] + 4 * brev_c(q, LR))] = v[k][q];
        }
    }
#else
#pragma unroll 1
    for (int k = 0; k < PER; ++k) {
        const int idx = swapbit2(t + k * M, sb);
        const int o = idx & 3;
        const int rest = idx >> 2;
        const int g = rest & ((1 << l2gr) - 1);
        const int wv = rest >> l2gr;
        const int base = low2_wbase(wv >> 1, wv & 1, l2n) + (g << (LR + 2)) + o;
        cplx v[R];
        if (DIR > 0) {
#pragma unroll
            for (int q = 0; q < R; ++q)
                v[q] = S[swz2(base + 4 * brev_c(q, LR))];
            up_twiddle_mul<R, true>(v, pt, 4, o);
            Dft<R, +1>::run(v);
#pragma unroll
            for (int n = 0; n < R; ++n)
                S[swz2(base + 4 * n)] = v[n];
        } else {
#pragma unroll
            for (int n = 0; n < R; ++n)
                v[n] = S[swz2(base + 4 * n)];
            Dft<R, -1>::run(v);
            up_twiddle_mul<R, false>(v, pt, 4, o);
#pragma unroll
            for (int q = 0; q < R; ++q)
                S[swz2(base + 4 * brev_c(q, LR))] = v[q];
        }
    }
#endif
}

// ---------------------------------------------------------------------------------------
// M stage: last inverse pass (radix R, stride s = N/R), coefficient fix-up, twist by
// w_2N^i / N, first forward pass -- all on one register set.
// ---------------------------------------------------------------------------------------
template <int LOG2M, int R>
DEV double low2_m_stage(cplx *S, const TwSet &tw, const Low2Tops *TTn, int t, int l2n, int sb, bool want_max)
{
    constexpr int M = 1 << LOG2M;
    constexpr int LR = Log2R<R>::value;
    constexpr int PER = 16 / R;
    const int l2s = l2n - LR;
    const int s = 1 << l2s;
    const double invN = 1.0 / (double)(1 << l2n);
    const cplx *pt = tw.base + tw.pass_off[l2n][LR];  // [q-1][o], o < s
    const cplx *tt = tw.base + tw.twist_off[l2n];     // w_2N^i
    double m2 = 0.0;
#if FNFTB_LOW2_BATCH >= 2
    if constexpr (PER > 1) {
        cplx v[PER][R];
        int bases[PER], os[PER], ps[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            int idx = t + k * M;
            if (l2s == 2)
                idx = swapbit2(idx, sb);
            os[k] = idx & (s - 1);
            const int wv = idx >> l2s;
            ps[k] = wv;
            bases[k] = low2_wbase(wv >> 1, wv & 1, l2n) + os[k];
#pragma unroll
            for (int q = 0; q < R; ++q)
                v[k][q] = S[swz2(bases[k] + (brev_c(q, LR) << l2s))];
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            up_twiddle_mul<R, true>(v[k], pt, s, os[k]);
            Dft<R, +1>::run(v[k]);
#pragma unroll
            for (int n = 0; n < R; ++n)
                v[k][n] = cscale(v[k][n], invN);
            if (os[k] == 0)
                v[k][0] = (ps[k] & 1) ? TTn[ps[k] >> 1].bb : TTn[ps[k] >> 1].ba;
            if (want_max) {
#pragma unroll
                for (int n = 0; n < R; ++n)
                    m2 = fmax(m2, cabs2(v[k][n]));
            }
            UpTwist<R, 0>::run(v[k], __ldg(&tt[os[k]]));
            Dft<R, -1>::run(v[k]);
            up_twiddle_mul<R, false>(v[k], pt, s, os[k]);
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
#pragma unroll
            for (int q = 0; q < R; ++q)
                S[swz2(bases[k] + (brev_c(q, LR) << l2s))] = v[k][q];
        }
        return m2;
    }
#endif
#pragma unroll 1
    for (int k = 0; k < PER; ++k) {
        int idx = t + k * M;
        if (l2s == 2)
            idx = swapbit2(idx, sb);
        const int o = idx & (s - 1);
        const int wv = idx >> l2s;
        const int p = wv >> 1, which = wv & 1;
        const int base = low2_wbase(p, which, l2n) + o;
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q)
            v[q] = S[swz2(base + (brev_c(q, LR) << l2s))];
        up_twiddle_mul<R, true>(v, pt, s, o);
        Dft<R, +1>::run(v);
        // v[n] = N * c[o + n*s]
#pragma unroll
        for (int n = 0; n < R; ++n)
            v[n] = cscale(v[n], invN);
        if (o == 0)
            v[0] = which ? TTn[p].bb : TTn[p].ba;
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

// spec_out: last forward pass (radix 4, stride 1) of the two workspace arrays of the single
// remaining pair, "- c_N", store the odd bins of the result spectrum
template <int LOG2M>
DEV void low2_final_stage(const cplx *S, const Low2Tops &Tn, int t, int l2n, cplx *gout)
{
    constexpr int M = 1 << LOG2M;
    const int N = 1 << l2n;
    const int items = N >> 1;  // 2 arrays * N/4 groups
#pragma unroll 1
    for (int it = t; it < items; it += M) {
        const int which = it >> (l2n - 2);
        const int g = it & ((N >> 2) - 1);
        const int ad = swz2(low2_wbase(0, which, l2n) + 4 * g);
        cplx v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            v[j] = S[ad ^ j];
        Dft<4, -1>::run(v);
        const cplx ct = which ? Tn.tb : Tn.ta;
        cplx *dst = gout + (size_t)which * (2 * N) + N + 4 * g;
        dst[0] = csub(v[0], ct);
        dst[1] = csub(v[2], ct);
        dst[2] = csub(v[1], ct);
        dst[3] = csub(v[3], ct);
    }
}

// last level: final inverse pass (radix 16), 1/N, write the coefficients; returns max |c|^2
template <int LOG2M>
DEV double low2_out_stage(const cplx *S, const TwSet &tw, const Low2Tops &Tn, int t, int l2n, cplx *out)
{
    constexpr int R = 16, LR = 4;
    const int l2s = l2n - LR;
    const int s = 1 << l2s;
    const int N = 1 << l2n;
    const double invN = 1.0 / (double)N;
    const int o = t & (s - 1);
    const int which = t >> l2s;  // M threads == 2 arrays * s items
    const int base = low2_wbase(0, which, l2n) + o;
    const cplx *pt = tw.base + tw.pass_off[l2n][LR];
    cplx v[R];
#pragma unroll
    for (int q = 0; q < R; ++q)
        v[q] = S[swz2(base + (brev_c(q, LR) << l2s))];
    up_twiddle_mul<R, true>(v, pt, s, o);
    Dft<R, +1>::run(v);
    cplx *dst = out + (size_t)which * (N + 1);
    double m2 = 0.0;
#pragma unroll
    for (int n = 0; n < R; ++n) {
        cplx c = cscale(v[n], invN);
        if (n == 0 && o == 0)
            c = which ? Tn.bb : Tn.ba;
        dst[o + (n << l2s)] = c;
        m2 = fmax(m2, cabs2(c));
    }
    if (o == 0) {
        const cplx ct = which ? Tn.tb : Tn.ta;
        dst[N] = ct;
        m2 = fmax(m2, cabs2(ct));
    }
    return m2;
}

// radix plan per operand length: P pass radix (0 = none), M stage radix, lane swap bit
DEV void low2_plan(int l2n, int *rp, int *rm, int *sbp, int *sbm)
{
    *sbp = 2;
    *sbm = 2;
    switch (l2n) {
    case 4: *rp = 0; *rm = 4; break;
    case 5: *rp = 0; *rm = 8; *sbm = 4; break;
    case 6: *rp = 0; *rm = 16; *sbm = 3; break;
    case 7: *rp = 4; *rm = 8; *sbp = 3; break;
    case 8: *rp = 4; *rm = 16; *sbp = 3; break;
    case 9: *rp = 8; *rm = 16; break;
    default: *rp = 16; *rm = 16; *sbp = 4; break;  // 1024
    }
}

// grid.x = B * (npad / S), S = M * 8 / DEG0, blockDim.x = M
template <int LOG2M, int DEG0>
__global__ void __launch_bounds__(1 << LOG2M, (LOG2M == 6) ? 6 : 3) k_tree_low2(const Low2Args a)
{
    constexpr int M = 1 << LOG2M;
    constexpr int LC = (DEG0 == 2) ? 2 : 3;  // log2(samples per thread)
    constexpr int S_ = M << LC;
    extern __shared__ double2 fnftb_smem2[];
    cplx *S = (cplx *)fnftb_smem2;                 // 32*M
    Low2Tops *TT0 = (Low2Tops *)(S + 32 * M);      // M/2
    Low2Tops *TT1 = TT0 + M / 2;                   // M/4
    double *red = (double *)(TT1 + M / 4);         // M/32 doubles
    const int t = threadIdx.x;
    const int nblk = a.npad / S_;
    const int s = blockIdx.x / nblk, blk = blockIdx.x % nblk;
    const double kap = (double)a.kappa;

    // ---- thread phase: degree-8 matrix of this thread's samples -----------------------
    Low2Tops mine;
    {
        cplx A[9], Bc[9];
        int err = 0;
        Low2Build<DEG0, LC>::run(a, s, blk * S_ + (t << LC), A, Bc, &err);
        if (err && a.status)
            a.status[s] = err;
        int ex = 0;
        if (a.normalize) {
            double m = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                m = fmax(m, fmax(fabs(A[i].x), fabs(A[i].y)));
                m = fmax(m, fmax(fabs(Bc[i].x), fabs(Bc[i].y)));
            }
            ex = rescale_exponent(m);
            if (ex != 0) {
                const double sc = ldexp(1.0, -ex);
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    A[i] = cscale(A[i], sc);
                    Bc[i] = cscale(Bc[i], sc);
                }
            }
            // W[s] += sum of the exponents of the CTA
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
                ex += __shfl_xor_sync(0xffffffffu, ex, off);
            if ((t & 31) == 0 && ex != 0)
                atomicAdd(&a.W[s], ex);
        }
        mine.ta = A[8];
        mine.tb = Bc[8];
        mine.ba = A[0];
        mine.bb = Bc[0];
        low2_front_fft16(A, S, t * 32);
        low2_front_fft16(Bc, S, t * 32 + 16);
    }
    // tops of the level-1 matrices: even lanes combine with their right neighbour
    {
        Low2Tops nb;
        nb.ta.x = __shfl_down_sync(0xffffffffu, mine.ta.x, 1);
        nb.ta.y = __shfl_down_sync(0xffffffffu, mine.ta.y, 1);
        nb.tb.x = __shfl_down_sync(0xffffffffu, mine.tb.x, 1);
        nb.tb.y = __shfl_down_sync(0xffffffffu, mine.tb.y, 1);
        nb.ba.x = __shfl_down_sync(0xffffffffu, mine.ba.x, 1);
        nb.ba.y = __shfl_down_sync(0xffffffffu, mine.ba.y, 1);
        nb.bb.x = __shfl_down_sync(0xffffffffu, mine.bb.x, 1);
        nb.bb.y = __shfl_down_sync(0xffffffffu, mine.bb.y, 1);
        if ((t & 1) == 0)
            TT0[t >> 1] = low2_pair_tops(mine, nb, kap, 1.0);
    }
    __syncthreads();

    // ---- levels: operand length 16, 32, ..., 8*M ---------------------------------------
    Low2Tops *TTc = TT1, *TTn = TT0;  // TTn: tops of the matrices being produced
    const size_t mo = (size_t)s * nblk + blk;
    cplx *gspec = a.spec_out ? a.out + mo * (size_t)(32 * M) : nullptr;  // [a: 16M][b: 16M]
    double m2 = 0.0;
#pragma unroll 1
    for (int L = 0; L < LOG2M; ++L) {
        const int l2n = 4 + L;
        const bool last = (L == LOG2M - 1);
        int rp, rm, sbp, sbm;
        low2_plan(l2n, &rp, &rm, &sbp, &sbm);
        if (L > 0 && t < (M >> (L + 1)))
            TTn[t] = low2_pair_tops(TTc[2 * t], TTc[2 * t + 1], kap, 1.0);
        low2_x_stage<LOG2M>(S, TTc, t, l2n, L == 0, last, kap, gspec);
        __syncthreads();
        switch (rp) {
        case 4: low2_p_pass<LOG2M, 4, +1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 8: low2_p_pass<LOG2M, 8, +1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 16:
            if constexpr (LOG2M > 6) {
                low2_p_pass<LOG2M, 16, +1>(S, a.tw, t, l2n, sbp);
                __syncthreads();
            }
            break;
        default: break;
        }
        if (last && !a.spec_out) {
            m2 = low2_out_stage<LOG2M>(S, a.tw, TTn[0], t, l2n, a.out + mo * 2 * ((8 * M) + 1));
            break;
        }
        const bool want_max = last;
        double mm;
        switch (rm) {
        case 4: mm = low2_m_stage<LOG2M, 4>(S, a.tw, TTn, t, l2n, sbm, want_max); break;
        case 8: mm = low2_m_stage<LOG2M, 8>(S, a.tw, TTn, t, l2n, sbm, want_max); break;
        default: mm = low2_m_stage<LOG2M, 16>(S, a.tw, TTn, t, l2n, sbm, want_max); break;
        }
        m2 = fmax(m2, mm);
        __syncthreads();
        switch (rp) {
        case 4: low2_p_pass<LOG2M, 4, -1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 8: low2_p_pass<LOG2M, 8, -1>(S, a.tw, t, l2n, sbp); __syncthreads(); break;
        case 16:
            if constexpr (LOG2M > 6) {
                low2_p_pass<LOG2M, 16, -1>(S, a.tw, t, l2n, sbp);
                __syncthreads();
            }
            break;
        default: break;
        }
        if (last) {  // spec_out: odd bins of the result
            low2_final_stage<LOG2M>(S, TTn[0], t, l2n, gspec);
            if (t == 0) {
                a.tt_out[mo] = TTn[0];
                m2 = fmax(m2, fmax(cabs2(TTn[0].ta), cabs2(TTn[0].tb)));
            }
            break;
        }
        Low2Tops *tmp = TTc;
        TTc = TTn;
        TTn = tmp;
    }
    // max |c| of the result for the lazy normalisation of the next level
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

static inline size_t low2_smem_bytes(int log2m)
{
    const size_t M = (size_t)1 << log2m;
    return sizeof(cplx) * 32 * M + sizeof(Low2Tops) * (M / 2 + M / 4) + sizeof(double) * 8;
}

// samples per CTA for (log2m, deg0)
static inline int low2_samples(int log2m, int deg0) { return (1 << log2m) * (8 / deg0); }

// The kernels are instantiated in ONE translation unit (k_tree_low2.cu); every other one sees the prototype.
#ifdef FNFTB_TU_LOW2
template <int LOG2M, int DEG0>
static inline int low2_launch_t(const Low2Args &a, cudaStream_t st)
{
    const size_t smem = low2_smem_bytes(LOG2M);
    const int e = fnftb_smem_optin((const void *)k_tree_low2<LOG2M, DEG0>, smem);
    if (e != 0)
        return e;
    const unsigned grid = (unsigned)a.B * (unsigned)(a.npad / low2_samples(LOG2M, DEG0));
    if (g_fnftb_profile_on)
        fnftb_profile_begin("tree_low2", st);
    k_tree_low2<LOG2M, DEG0><<<grid, 1 << LOG2M, smem, st>>>(a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(st);
    ++g_fnftb_launch_count;
    return (int)cudaGetLastError();
}

int low2_launch(const Low2Args &a, int log2m, int deg0, cudaStream_t st)
{
    if (log2m == 6)
        return deg0 == 2 ? low2_launch_t<6, 2>(a, st) : low2_launch_t<6, 1>(a, st);
    return deg0 == 2 ? low2_launch_t<7, 2>(a, st) : low2_launch_t<7, 1>(a, st);
}
#else
int low2_launch(const Low2Args &a, int log2m, int deg0, cudaStream_t st);
#endif
#endif  // !FNFTB_EMUL

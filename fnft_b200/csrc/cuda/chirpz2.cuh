// fnft_b200 -- fast path of the batched chirp-z evaluation (fnft__poly_chirpz.c:33-105)
// for Bluestein lengths L = R * 4096, R in {2, 3, 4, 6, 8, 12, 16} (the smallest that holds deg + M: BASELINE
// config 2 runs L = 12 * 4096 = 49152 like the reference's next_fast_size, fnft__poly_chirpz.c:52-56, instead
// of the next power of two 65536): same algorithm and tables as
// chirpz_kernels.cuh, but the two length-L transforms are organised like the tree kernels
// (tree_up.cuh): bit-reversed DIF forward / DIT inverse, pass-major twiddle tables, and
//
//   cz2_cols_fwd : y_n = p[deg-n]*tab_y[n] on the fly, radix-R DIF pass across the rows
//                  (stride 4096), streaming, one thread per column
//   cz2_rows     : per row of 4096: [global -> radix-16 (s=256)] [radix-16 (s=16)]
//                  [radix-16 (s=1) * FFT(v) * inverse radix-16 (s=1), in registers]
//                  [inverse radix-16 (s=16)] [inverse radix-16 (s=256) -> global]
//   cz2_cols_inv : radix-R DIT pass across the rows, only outputs m < M, * W^(m^2/2)/L,
//                  continuous-spectrum epilogue (src/fnft_nsev.c:846-876,
//                  src/fnft_kdvv.c:186-203) for both polynomials of a signal at once
//
// FFT(v) is stored pre-permuted ([row][q][g], value at row position 16*g + brev4(q)) so that
// the register-fused stage reads it with unit stride across the lanes.
#pragma once
#ifndef FNFTB_EMUL
#include "chirpz_kernels.cuh"
#include "tree_up.cuh"

#define FNFTB_CZ2_ROW_L2 12

// first-row-only tree result as polynomial source (instead of the finalised transfer matrix):
// level buffer layout of tree_kernels.cuh with E = 2, one matrix per signal of degree d_full
struct Cz2SymSrc {
    const cplx *lev;   // NULL: read c.tm
    const double *mx;  // [B] max|coeff| -> pending power-of-two scale
    int *W;            // [B] the exponent of that scale is added here (what blk_tree_final does)
    int d_full, kappa, normalize;
    // fused = 1: the coefficients do not exist yet -- `up` describes the pending column pass of the last tree level
    // (tree_up.cuh: k_up_cols, last = 1) and k_up_cols_cz computes them on the way into the first chirp-z stage
    int fused;
    UpArgs up;
};

struct Cz2Args {
    CzArgs c;     // polynomials, tables (tab_y, tab_out, tab_ph), epilogue description
    Cz2SymSrc src;
    TwSet tw;
    cplx *vperm;  // [L] permuted FFT(v)
    int l2L;      // log2 L when L is a power of two (pass-major twiddle table of the column pass)
    int R, L;     // rows, L = R * 4096
    int gen_v;    // cols_fwd: generate the chirp filter; rows: forward half only -> vperm
    int pf;       // L2 prefetch distance in CTAs (tree_up.cuh: l2_prefetch), 0 = off
};

// number of rows of 4096 for a Bluestein length >= need (0: not on the fast path)
static inline int cz2_rows_for(size_t need)
{
    static const int cand[7] = {2, 3, 4, 6, 8, 12, 16};
    if (need <= ((size_t)1 << FNFTB_CZ2_ROW_L2))
        return 0;  // short transforms take the general path
    for (int i = 0; i < 7; ++i)
        if (((size_t)cand[i] << FNFTB_CZ2_ROW_L2) >= need)
            return cand[i];
    return 0;
}

// grid.x * 256 threads = narr * 4096 ; narr = B*npoly (or 1 for gen_v)
static inline bool cz2_supported(int deg, int M)
{
    return cz2_rows_for((size_t)deg + (size_t)M) != 0 && FNFTB_CZ2_ROW_L2 + 4 <= FNFTB_TW_MAXL;
}

// can the last column pass of the spectrum-carry tree (radix RT = 2^(l2n - 12) across rows of 4096) be fused with the
// forward column pass of the chirp-z (RC rows)?  Needs an unpadded product (deg == d_full, i.e. D a power of two) and
// one of the instantiated (RT, RC) pairs
static inline bool cz2_fused_cols_supported(const UpArgs &u, int deg, int M, int d_full)
{
    if (deg != d_full || u.l2row != FNFTB_CZ2_ROW_L2 || deg != (1 << u.l2n))
        return false;
    const int RT = 1 << (u.l2n - u.l2row), RC = cz2_rows_for((size_t)deg + (size_t)M);
    return (RT == 2 && (RC == 3 || RC == 4)) || (RT == 4 && (RC == 6 || RC == 8)) ||
           (RT == 8 && (RC == 12 || RC == 16));
}

// kernels instantiated in k_chirpz2.cu only
#ifdef FNFTB_TU_CZ2
// ---- column transforms of length R = 3 * P (P = 1, 2, 4) next to the power-of-two Dft<R> ---------
template <int R>
struct Cz2Radix {
    static constexpr bool pow2 = (R & (R - 1)) == 0;
    static constexpr int LR = Log2R<R>::value;
    // row that holds element q of the column transform's digit-permuted side
    DEV static constexpr int row(int q) { return pow2 ? brev_c(q, LR) : q; }
};

// v *= exp(DIR * 2 pi i K / 12)
template <int K, int DIR>
DEV cplx mul_root12(cplx v)
{
    constexpr int k = ((K % 12) + 12) % 12;
    constexpr double h = 0.5, c30 = 0.8660254037844386;
    constexpr double cs[12] = {1.0, c30, h, 0.0, -h, -c30, -1.0, -c30, -h, 0.0, h, c30};
    constexpr double sn[12] = {0.0, h, c30, 1.0, c30, h, 0.0, -h, -c30, -1.0, -c30, -h};
    if constexpr (k == 0) {
        return v;
    } else if constexpr (k == 6) {
        return cneg(v);
    } else if constexpr (k == 3) {
        return (DIR > 0) ? cmuli(v) : cmulmi(v);
    } else if constexpr (k == 9) {
        return (DIR > 0) ? cmulmi(v) : cmuli(v);
    } else {
        constexpr double c = cs[k];
        constexpr double s = (DIR > 0) ? sn[k] : -sn[k];
        return make_cplx(v.x * c - v.y * s, v.x * s + v.y * c);
    }
}

template <int DIR>
DEV void dft3(cplx &a, cplx &b, cplx &c)
{
    constexpr double s60 = (DIR > 0) ? 0.8660254037844386 : -0.8660254037844386;
    const cplx t1 = cadd(b, c);
    const cplx t2 = make_cplx(a.x - 0.5 * t1.x, a.y - 0.5 * t1.y);
    const cplx d = csub(b, c);
    const cplx t3 = make_cplx(-s60 * d.y, s60 * d.x);  // i * s60 * (b - c)
    a = cadd(a, t1);
    b = cadd(t2, t3);
    c = csub(t2, t3);
}

template <int R, int DIR, int N2_, int K1>
struct Cz2Tw3 {  // y[N2_][K1] *= w_R^(N2_ * K1)
    DEV static void run(cplx (*y)[R / 3])
    {
        if constexpr (K1 < R / 3) {
            y[N2_][K1] = mul_root12<N2_ * K1 * (12 / R), DIR>(y[N2_][K1]);
            Cz2Tw3<R, DIR, N2_, K1 + 1>::run(y);
        }
    }
};

// natural order in and out, X[k] = sum_n x[n] exp(DIR 2 pi i n k / R)
template <int R, int DIR>
struct DftAny {
    DEV static void run(cplx *v)
    {
        if constexpr (Cz2Radix<R>::pow2) {
            Dft<R, DIR>::run(v);
        } else {
            // n = 3 n1 + n2, k = k1 + P k2:  w^(nk) = w_P^(n1 k1) * w_R^(n2 k1) * w_3^(n2 k2)
            constexpr int P = R / 3;
            cplx y[3][P];
#pragma unroll
            for (int n2 = 0; n2 < 3; ++n2) {
#pragma unroll
                for (int n1 = 0; n1 < P; ++n1)
                    y[n2][n1] = v[3 * n1 + n2];
                Dft<P, DIR>::run(y[n2]);
            }
            Cz2Tw3<R, DIR, 1, 0>::run(y);
            Cz2Tw3<R, DIR, 2, 0>::run(y);
#pragma unroll
            for (int k1 = 0; k1 < P; ++k1) {
                dft3<DIR>(y[0][k1], y[1][k1], y[2][k1]);
                v[k1] = y[0][k1];
                v[k1 + P] = y[1][k1];
                v[k1 + 2 * P] = y[2][k1];
            }
        }
    }
};

// v[q] *= w1^q, q < R (powers by squarings / products of depth <= 4)
template <int R>
DEV void cz2_pow_mul(cplx *v, cplx w1)
{
    cplx w[R < 2 ? 2 : R];
    w[1] = w1;
#pragma unroll
    for (int q = 2; q < R; ++q)
        w[q] = (q & 1) ? cmul(w[q - 1], w1) : csq(w[q / 2]);
#pragma unroll
    for (int q = 1; q < R; ++q)
        v[q] = cmul(v[q], w[q]);
}

// twiddle of the four-step split: element q of column o gets w_L^(o q) (CONJ: its conjugate)
template <int R, bool CONJ>
DEV void cz2_col_twiddle(cplx *v, const Cz2Args &a, int o)
{
    constexpr int N2 = 1 << FNFTB_CZ2_ROW_L2;
    if constexpr (Cz2Radix<R>::pow2) {
        const cplx *pt = a.tw.base + a.tw.pass_off[a.l2L][Cz2Radix<R>::LR];
        up_twiddle_mul<R, CONJ>(v, pt, N2, o);
    } else {
        double sn, cs;
        sincospi(2.0 * (double)o / (double)a.L, &sn, &cs);
        cz2_pow_mul<R>(v, make_cplx(cs, CONJ ? sn : -sn));
    }
}

template <int R>
__global__ void __launch_bounds__(256, 3) k_cz2_cols_fwd(const Cz2Args a)
{
    constexpr int N2 = 1 << FNFTB_CZ2_ROW_L2;
    const CzArgs &c = a.c;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int o = (int)(gid & (N2 - 1));
    const size_t arr = (size_t)(gid >> FNFTB_CZ2_ROW_L2);
    const int Np = c.deg + 1;
    const int L = a.L;
    cplx v[R];
    if (!a.gen_v && a.src.lev) {
        // H11[deg-n] = scale * a[deg-n];  H21[deg-n] = -kappa * scale * conj(b[shift + n])
        // (blk_tree_final: T21 = -kappa * T12#, second column shifted by the padding)
        const int j = (int)(arr & 1);
        const size_t s = arr >> 1;
        int ex = 0;
        double scale = 1.0;
        if (a.src.normalize) {
            ex = rescale_exponent(a.src.mx[s]);
            scale = ldexp(1.0, -ex);
        }
        if (j == 0 && o == 0 && ex != 0)
            a.src.W[s] += ex;  // this signal's tree kernels have all finished
        const cplx *pl = a.src.lev + (s * 2 + j) * (size_t)(a.src.d_full + 1);
        const int shift = a.src.d_full - c.deg;
        const double f = j ? -(double)a.src.kappa * scale : scale;
        if (a.pf > 0 && threadIdx.x < R) {  // coefficient segments of the CTA `pf` places ahead
            const size_t bp = (size_t)blockIdx.x + (size_t)a.pf;
            const long long g0 = (long long)bp * blockDim.x;
            const int n_lo = (int)(g0 & (N2 - 1)) + (int)threadIdx.x * N2;
            int n_hi = n_lo + (int)blockDim.x;
            n_hi = n_hi < Np ? n_hi : Np;
            if (bp < gridDim.x && n_lo < n_hi) {
                const size_t arrp = (size_t)(g0 >> FNFTB_CZ2_ROW_L2);
                const cplx *pp = a.src.lev + arrp * (size_t)(a.src.d_full + 1);
                const cplx *first = (arrp & 1) ? pp + shift + n_lo : pp + c.deg - (n_hi - 1);
                l2_prefetch(first, (unsigned)(n_hi - n_lo) * (unsigned)sizeof(cplx));
            }
        }
#pragma unroll
        for (int n1 = 0; n1 < R; ++n1) {
            const int n = o + n1 * N2;
            if (n < Np) {
                cplx x = j ? pl[shift + n] : pl[c.deg - n];
                x = make_cplx(x.x * f, j ? -x.y * f : x.y * f);
                v[n1] = cmul(x, __ldg(&c.tab_y[n]));
            } else {
                v[n1] = czero();
            }
        }
    } else if (!a.gen_v) {
        const int j = (int)(arr % c.npoly);
        const size_t s = arr / c.npoly;
        const cplx *p = c.tm + s * c.tm_sstride + (size_t)c.ent[j] * Np;
#pragma unroll
        for (int n1 = 0; n1 < R; ++n1) {
            const int n = o + n1 * N2;
            v[n1] = (n < Np) ? cmul(p[c.deg - n], __ldg(&c.tab_y[n])) : czero();
        }
    } else {
#pragma unroll
        for (int n1 = 0; n1 < R; ++n1) {
            const int n = o + n1 * N2;
            // fnft__poly_chirpz.c:76-82
            double dn = -1.0;
            if (n < c.M)
                dn = (double)n;
            else if (n > L - Np)
                dn = (double)(L - n);
            v[n1] = (dn >= 0.0) ? (c.dft_n > 0 ? chirp_dft((long long)dn, c.dft_n, c.lwi < 0.0 ? 1.0 : -1.0)
                                               : chirp_factor(-c.lwr * (0.5 * dn * dn), -c.lwi, 0.5 * dn * dn, 0.0, 0.0))
                                : czero();
        }
    }
    DftAny<R, -1>::run(v);
    cz2_col_twiddle<R, false>(v, a, o);
    cplx *dst = c.ybuf + arr * (size_t)L;
#pragma unroll
    for (int q = 0; q < R; ++q)
        dst[((size_t)Cz2Radix<R>::row(q) << FNFTB_CZ2_ROW_L2) + o] = v[q];
}

// FUSED (round 2): last column pass of the product tree + first chirp-z stage.  k_up_cols<RT> (last level) turns
// column o of the row-split workspace into the coefficients g[o + n*4096], n < RT (+ g[N] from the tops for o = 0);
// k_cz2_cols_fwd<RC> needs, for chirp column o', the coefficients g[N - n] (polynomial a) or g[n] (polynomial b) with
// n = o' + n1*4096: for b that is the thread's own column (o' = o), for a it is column o' = (4096 - o) mod 4096 with the
// rows in reverse order -- either way every value a thread needs it has just computed.  The coefficients are never
// written (2 MB per signal less DRAM traffic each way).  No scaling by max|c| here: the max over the whole matrix is
// not known before all CTAs have finished, powers of two commute with everything, and rho = b / a does not see it; the
// exponent W[s] keeps describing the values that flow on (the epilogue multiplies a and b by 2^W).
template <int RT, int RC>
__global__ void __launch_bounds__(256, 3) k_up_cols_cz(const Cz2Args a)
{
    constexpr int N2 = 1 << FNFTB_CZ2_ROW_L2;
    constexpr int LR = Log2R<RT>::value;
    const UpArgs &u = a.src.up;
    const CzArgs &c = a.c;
    const int l2n = u.l2n, l2row = l2n - LR;
    const int N = 1 << l2n;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int o = (int)(gid & (N2 - 1));
    const size_t arr = (size_t)(gid >> l2row);  // = 2 * signal + which (one pair per signal on the last level)
    const int which = (int)(arr & 1);
    const size_t sp = arr >> 1;
    const Low2Tops Tn = ((const Low2Tops *)u.tt_out)[sp];
    const cplx *w = u.ws + arr * (size_t)N;
    const cplx *pt = u.tw.base + u.tw.pass_off[l2n][LR];
    const double invN = 1.0 / (double)N;
    if (u.pf > 0 && threadIdx.x < RT) {  // the RT column segments of the CTA `pf` places ahead
        const size_t bp = (size_t)blockIdx.x + (size_t)u.pf;
        if (bp < gridDim.x) {
            const long long g0 = (long long)bp * blockDim.x;
            const cplx *wp = u.ws + (size_t)(g0 >> l2row) * (size_t)N + (size_t)(g0 & ((1 << l2row) - 1));
            l2_prefetch(wp + ((size_t)brev_c((int)threadIdx.x, LR) << l2row), blockDim.x * (unsigned)sizeof(cplx));
        }
    }
    cplx g[RT];
#pragma unroll
    for (int q = 0; q < RT; ++q)
        g[q] = w[o + ((size_t)brev_c(q, LR) << l2row)];
    up_twiddle_mul<RT, true>(g, pt, 1 << l2row, o);
    Dft<RT, +1>::run(g);
#pragma unroll
    for (int n = 0; n < RT; ++n)
        g[n] = cscale(g[n], invN);
    if (o == 0)
        g[0] = up_bot(Tn, which);
    const cplx gtop = up_top(Tn, which);  // g[N], used by the column-0 thread only
    // ---- first chirp-z stage: y_n = x_n * tab_y[n], radix-RC DIF pass across the rows
    const int Np = c.deg + 1;
    const int oc = which ? o : ((N2 - o) & (N2 - 1));  // chirp column this thread feeds
    const double fb = -(double)a.src.kappa;
    cplx v[RC];
#pragma unroll
    for (int n1 = 0; n1 < RC; ++n1) {
        const int n = oc + n1 * N2;
        cplx x = czero();
        if (n < Np) {
            if (which) {  // H21[deg - n] = -kappa conj(b[n])
                const cplx b = (n1 < RT) ? g[n1 < RT ? n1 : 0] : gtop;  // n1 == RT only occurs for o == 0 (n == N)
                x = make_cplx(b.x * fb, -b.y * fb);
            } else {      // H11[deg - n] = a[N - n]
                if (o == 0)
                    x = (n1 == 0) ? gtop : g[(RT - n1) >= 0 && (RT - n1) < RT ? RT - n1 : 0];
                else
                    x = g[(RT - 1 - n1) >= 0 ? RT - 1 - n1 : 0];
            }
            x = cmul(x, __ldg(&c.tab_y[n]));
        }
        v[n1] = x;
    }
    DftAny<RC, -1>::run(v);
    cz2_col_twiddle<RC, false>(v, a, oc);
    cplx *dst = c.ybuf + arr * (size_t)a.L;
#pragma unroll
    for (int q = 0; q < RC; ++q)
        dst[((size_t)Cz2Radix<RC>::row(q) << FNFTB_CZ2_ROW_L2) + oc] = v[q];
}

// The WHOLE last tree level + first chirp-z stage as a cluster of RT CTAs (round 2).  CTA h of a cluster holds row h
// (4096 positions, 64 KiB) of one (signal, polynomial): X stage and the inverse passes inside the row like k_up_rows_a,
// then -- instead of writing the row to the workspace for k_up_cols -- a cluster barrier and the column stage of
// k_up_cols_cz straight from the RT shared memories (distributed shared memory reads), so that between the operands of
// the last pair product and the first chirp-z buffer nothing touches DRAM.
// grid.x = B * 2 * RT, cluster (RT, 1, 1), 128 threads
template <int RT, int RC>
__global__ void __launch_bounds__(128, 3) k_up_last_cluster_cz(const Cz2Args a)
{
    constexpr int N2 = 1 << FNFTB_CZ2_ROW_L2, NT = 128;
    constexpr int LR = Log2R<RT>::value;
    extern __shared__ double2 fnftb_smem_cz2[];
    cplx *S = (cplx *)fnftb_smem_cz2;
    const UpArgs &u = a.src.up;
    const CzArgs &c = a.c;
    const int tid = threadIdx.x;
    const unsigned h = up_cluster_rank();
    const unsigned item = blockIdx.x / RT;   // = 2 * signal + which (one pair per signal on the last level)
    const int which = (int)(item & 1);
    const size_t sp = item >> 1;
    const int l2n = u.l2n;
    const int N = 1 << l2n;
    if (u.pf_rows > 0 && h == 0) {
        const size_t ip = (size_t)item + (size_t)u.pf_rows;
        if (ip < (gridDim.x / RT) && (ip & 1) == 0)
            up_prefetch_operands<true>(u, ip >> 1, 0, 0, N, tid);
    }
    const UpPair<true> P = up_pair_setup<true>(u, sp, (int)sp, which == 0 && h == 0);
    up_x_stage<16, true>(u, sp, which, (int)h << FNFTB_CZ2_ROW_L2, 0, N2, (h >= RT / 2) ? -1.0 : 1.0, P, S, tid, NT);
    __syncthreads();
    up_row_passes<+1, FNFTB_UP_ROW_L2, true>(S, u.tw, tid, NT);
    up_cluster_sync();
    cplx *Sr[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
        Sr[r] = up_cluster_map(S, (unsigned)r);
    const cplx *pt = u.tw.base + u.tw.pass_off[l2n][LR];
    const double invN = 1.0 / (double)N;
    const int Np = c.deg + 1;
    const double fb = -(double)a.src.kappa;
    const size_t arr = item;
    cplx *dst = c.ybuf + arr * (size_t)a.L;
#pragma unroll 1
    for (int o = (int)h * (N2 / RT) + tid; o < ((int)h + 1) * (N2 / RT); o += NT) {
        cplx g[RT];
#pragma unroll
        for (int q = 0; q < RT; ++q)
            g[q] = Sr[brev_c(q, LR)][swz2(o)];
        up_twiddle_mul<RT, true>(g, pt, N2, o);
        Dft<RT, +1>::run(g);
#pragma unroll
        for (int n = 0; n < RT; ++n)
            g[n] = cscale(g[n], invN);
        if (o == 0)
            g[0] = up_bot(P.Tn, which);
        const cplx gtop = up_top(P.Tn, which);
        const int oc = which ? o : ((N2 - o) & (N2 - 1));
        cplx v[RC];
#pragma unroll
        for (int n1 = 0; n1 < RC; ++n1) {
            const int n = oc + n1 * N2;
            cplx x = czero();
            if (n < Np) {
                if (which) {
                    const cplx b = (n1 < RT) ? g[n1 < RT ? n1 : 0] : gtop;
                    x = make_cplx(b.x * fb, -b.y * fb);
                } else {
                    if (o == 0)
                        x = (n1 == 0) ? gtop : g[(RT - n1) >= 0 && (RT - n1) < RT ? RT - n1 : 0];
                    else
                        x = g[(RT - 1 - n1) >= 0 ? RT - 1 - n1 : 0];
                }
                x = cmul(x, __ldg(&c.tab_y[n]));
            }
            v[n1] = x;
        }
        DftAny<RC, -1>::run(v);
        cz2_col_twiddle<RC, false>(v, a, oc);
#pragma unroll
        for (int q = 0; q < RC; ++q)
            dst[((size_t)Cz2Radix<RC>::row(q) << FNFTB_CZ2_ROW_L2) + oc] = v[q];
    }
    up_cluster_sync();  // nobody leaves while a partner may still read its shared memory
}

// grid.x = narr * R rows, 128 threads, 64 KiB shared memory
__global__ void __launch_bounds__(128, 3) k_cz2_rows(const Cz2Args a)
{
    constexpr int N2 = 1 << FNFTB_CZ2_ROW_L2;
    constexpr int NT = 128;
    extern __shared__ double2 fnftb_smem_cz2[];
    cplx *S = (cplx *)fnftb_smem_cz2;
    const int tid = threadIdx.x;
    const int row = (int)(blockIdx.x % (unsigned)a.R);
    cplx *g = a.c.ybuf + (size_t)blockIdx.x * N2;  // rows are contiguous: arr*L + row*N2
    if (a.pf > 0 && (size_t)blockIdx.x + (size_t)a.pf < gridDim.x)
        l2_prefetch_span(g + (size_t)a.pf * N2, N2, tid, 0);
    const cplx *pt12 = a.tw.base + a.tw.pass_off[12][4];  // len 4096, radix 16, s = 256
    // global -> forward radix-16 at stride 256 -> shared
#pragma unroll 1
    for (int o = tid; o < 256; o += NT) {
        cplx v[16];
#pragma unroll
        for (int n = 0; n < 16; ++n)
            v[n] = g[o + (n << 8)];
        Dft<16, -1>::run(v);
        up_twiddle_mul<16, false>(v, pt12, 256, o);
#pragma unroll
        for (int q = 0; q < 16; ++q)
            S[swz2(o + (brev_c(q, 4) << 8))] = v[q];
    }
    __syncthreads();
    up_p_pass<16, -1>(S, N2, 4, a.tw, tid, NT);
    __syncthreads();
    // stride-1 forward pass, * FFT(v), stride-1 inverse pass (registers)
    const cplx *vp = a.vperm + ((size_t)row << FNFTB_CZ2_ROW_L2);
#pragma unroll 1
    for (int gi = tid; gi < 256; gi += NT) {
        cplx v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            v[j] = S[swz2(16 * gi + j)];
        Dft<16, -1>::run(v);
        if (a.gen_v) {
            cplx *vo = a.vperm + ((size_t)row << FNFTB_CZ2_ROW_L2);
#pragma unroll
            for (int q = 0; q < 16; ++q)
                vo[(q << 8) + gi] = v[q];
            continue;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q)
            v[q] = cmul(v[q], __ldg(&vp[(q << 8) + gi]));
        Dft<16, +1>::run(v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            S[swz2(16 * gi + j)] = v[j];
    }
    if (a.gen_v)
        return;
    __syncthreads();
    up_p_pass<16, +1>(S, N2, 4, a.tw, tid, NT);
    __syncthreads();
    // shared -> inverse radix-16 at stride 256 -> global
#pragma unroll 1
    for (int o = tid; o < 256; o += NT) {
        cplx v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q)
            v[q] = S[swz2(o + (brev_c(q, 4) << 8))];
        up_twiddle_mul<16, true>(v, pt12, 256, o);
        Dft<16, +1>::run(v);
#pragma unroll
        for (int n = 0; n < 16; ++n)
            g[o + (n << 8)] = v[n];
    }
}

// continuous-spectrum epilogue of output point m (both polynomials evaluated): H0, H1 already
// carry the factor W^(m^2/2)/L
DEV void cz2_epilogue(const CzArgs &c, size_t s, int m, cplx H0, cplx H1, cplx *out)
{
    if (c.mode == FNFTB_CZ_RAW) {
        const size_t js = c.out_jstride ? c.out_jstride : (size_t)c.M;
        out[m] = H0;
        if (c.npoly > 1)
            out[js + m] = H1;
    } else if (c.mode == FNFTB_CZ_NSEV) {
        // src/fnft_nsev.c:846-876; H0 = H11 (a-poly), H1 = H21 (b-poly)
        size_t off = 0;
        if (c.cstype == 0 || c.cstype == 2) {
            if (H0.x == 0.0 && H0.y == 0.0) {
                if (c.status)
                    c.status[s] = 3;
                out[m] = make_cplx(NAN, NAN);
            } else {
                out[m] = cdiv(cmul(H1, __ldg(&c.tab_ph[m])), H0);
            }
            off = c.M;
        }
        if (c.cstype == 1 || c.cstype == 2) {
            const double scale = ldexp(1.0, c.W ? c.W[s] : 0);
            out[off + m] = cmul(cscale(H0, scale), __ldg(&c.tab_ph[c.M + m]));
            out[off + c.M + m] = cmul(cscale(H1, scale), __ldg(&c.tab_ph[2 * (size_t)c.M + m]));
        }
    } else {
        // src/fnft_kdvv.c:186-203; H0 = H12, H1 = H22, xi grid negated
        const double xi = -c.xi0 - (double)m * c.eps_xi;
        cplx h12 = H0;
        if (c.kdv_sqrtz != 0.0)
            h12 = cdiv(h12, __ldg(&c.tab_ph[c.M + m]));
        const cplx num = cmul(__ldg(&c.tab_ph[m]), h12);
        const cplx den = make_cplx(-2.0 * xi * H1.y - h12.x, 2.0 * xi * H1.x - h12.y);
        out[m] = cdiv(num, den);
    }
}

// grid.x * 256 threads = B * 4096
template <int R>
__global__ void __launch_bounds__(256) k_cz2_cols_inv(const Cz2Args a)
{
    constexpr int N2 = 1 << FNFTB_CZ2_ROW_L2;
    const CzArgs &c = a.c;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int o = (int)(gid & (N2 - 1));
    const size_t s = (size_t)(gid >> FNFTB_CZ2_ROW_L2);
    if (o >= c.M)
        return;
    const int L = a.L;
    cplx H[2][R];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (j < c.npoly) {
            const cplx *src = c.ybuf + (s * c.npoly + j) * (size_t)L;
#pragma unroll
            for (int q = 0; q < R; ++q)
                H[j][q] = src[((size_t)Cz2Radix<R>::row(q) << FNFTB_CZ2_ROW_L2) + o];
            cz2_col_twiddle<R, true>(H[j], a, o);
            DftAny<R, +1>::run(H[j]);
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q)
                H[j][q] = czero();
        }
    }
    cplx *out = c.out + s * c.out_sstride;
#pragma unroll
    for (int n = 0; n < R; ++n) {
        const int m = o + n * N2;
        if (m >= c.M)
            break;
        const cplx ch = __ldg(&c.tab_out[m]);
        cz2_epilogue(c, s, m, cmul(H[0][n], ch), cmul(H[1][n], ch), out);
    }
}

// exp(2 pi i k / 16), folded to constants once the callers' loops are unrolled
DEV cplx unit_root16(int k)
{
    const double c1 = 0.9238795325112867, s1 = 0.3826834323650898, h = 0.7071067811865476;
    switch (k & 15) {
    case 0: return make_cplx(1.0, 0.0);
    case 1: return make_cplx(c1, s1);
    case 2: return make_cplx(h, h);
    case 3: return make_cplx(s1, c1);
    case 4: return make_cplx(0.0, 1.0);
    case 5: return make_cplx(-s1, c1);
    case 6: return make_cplx(-h, h);
    case 7: return make_cplx(-c1, s1);
    case 8: return make_cplx(-1.0, 0.0);
    case 9: return make_cplx(-c1, -s1);
    case 10: return make_cplx(-h, -h);
    case 11: return make_cplx(-s1, -c1);
    case 12: return make_cplx(0.0, -1.0);
    case 13: return make_cplx(s1, -c1);
    case 14: return make_cplx(h, -h);
    default: return make_cplx(c1, -s1);
    }
}

// exp(2 pi i k / 12)
DEV cplx unit_root12(int k)
{
    const double c30 = 0.8660254037844386, h = 0.5;
    switch (k % 12) {
    case 0: return make_cplx(1.0, 0.0);
    case 1: return make_cplx(c30, h);
    case 2: return make_cplx(h, c30);
    case 3: return make_cplx(0.0, 1.0);
    case 4: return make_cplx(-h, c30);
    case 5: return make_cplx(-c30, h);
    case 6: return make_cplx(-1.0, 0.0);
    case 7: return make_cplx(-c30, -h);
    case 8: return make_cplx(-h, -c30);
    case 9: return make_cplx(0.0, -1.0);
    case 10: return make_cplx(h, -c30);
    default: return make_cplx(c30, -h);
    }
}

// Output-pruned variant: only the rows n < NOUT (m = o + n*4096 < M) of the radix-R inverse column
// pass are wanted -- config 2 needs 4 of 16.  With q = s + S*t (S = R/NOUT):
//   X[n] = sum_s w_R^(s n) * ( sum_t x[s + S t] w_NOUT^(t n) ),      n < NOUT,
// i.e. S streamed NOUT-point transforms accumulated into NOUT values per polynomial: a fraction of
// the registers of the full butterfly (the kernel is latency bound: 174 registers, 8 warps per SM),
// so three times the warps are resident.
template <int R, int NOUT>
__global__ void __launch_bounds__(256, 3) k_cz2_cols_inv_p(const Cz2Args a)
{
    constexpr int N2 = 1 << FNFTB_CZ2_ROW_L2;
    constexpr int S = R / NOUT;
    static_assert(S * NOUT == R, "NOUT must divide R");
    const CzArgs &c = a.c;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int o = (int)(gid & (N2 - 1));
    const size_t s = (size_t)(gid >> FNFTB_CZ2_ROW_L2);
    const int L = a.L;
    if (a.pf > 0 && (int)threadIdx.x < R * c.npoly) {  // the column segments of the CTA `pf` places ahead
        const size_t bp = (size_t)blockIdx.x + (size_t)a.pf;
        const long long g0 = (long long)bp * blockDim.x;
        const int o0 = (int)(g0 & (N2 - 1));
        if (bp < gridDim.x && o0 < c.M) {
            const size_t sp = (size_t)(g0 >> FNFTB_CZ2_ROW_L2);
            const int j = (int)threadIdx.x / R, row = (int)threadIdx.x % R;
            l2_prefetch(c.ybuf + (sp * c.npoly + j) * (size_t)L + ((size_t)row << FNFTB_CZ2_ROW_L2) + o0,
                        blockDim.x * (unsigned)sizeof(cplx));
        }
    }
    if (o >= c.M)
        return;
    cplx w1;  // element q carries w1^q = conj(w_L^(o q)) (cz2_col_twiddle<R, true>)
    if constexpr (Cz2Radix<R>::pow2) {
        const cplx *pt = a.tw.base + a.tw.pass_off[a.l2L][Cz2Radix<R>::LR];
        w1 = cconj(__ldg(&pt[o]));
    } else {
        double sn, cs;
        sincospi(2.0 * (double)o / (double)L, &sn, &cs);
        w1 = make_cplx(cs, sn);
    }
    cplx wS = w1;
#pragma unroll
    for (int i = 1; i < S; ++i)
        wS = cmul(wS, w1);
    cplx acc[2][NOUT];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int n = 0; n < NOUT; ++n)
            acc[j][n] = czero();
        if (j < c.npoly) {
            const cplx *src = c.ybuf + (s * c.npoly + j) * (size_t)L;
            cplx ws = make_cplx(1.0, 0.0);
#pragma unroll
            for (int sI = 0; sI < S; ++sI) {
                cplx y[NOUT];
#pragma unroll
                for (int t = 0; t < NOUT; ++t)
                    y[t] = src[((size_t)Cz2Radix<R>::row(sI + S * t) << FNFTB_CZ2_ROW_L2) + o];
                cplx tw = ws;
#pragma unroll
                for (int t = 0; t < NOUT; ++t) {
                    if (sI + t > 0)
                        y[t] = cmul(y[t], tw);
                    tw = cmul(tw, wS);
                }
                DftAny<NOUT, +1>::run(y);
#pragma unroll
                for (int n = 0; n < NOUT; ++n) {
                    // w_R^(sI n): a 16th root of unity for R | 16, a 12th root for R | 12
                    const int k = Cz2Radix<R>::pow2 ? ((sI * n * (16 / (Cz2Radix<R>::pow2 ? R : 16))) & 15)
                                                    : ((sI * n * (12 / (Cz2Radix<R>::pow2 ? 12 : R))) % 12);
                    if (k == 0) {
                        acc[j][n] = cadd(acc[j][n], y[n]);
                    } else {
                        cfma(acc[j][n], y[n], Cz2Radix<R>::pow2 ? unit_root16(k) : unit_root12(k));
                    }
                }
                ws = cmul(ws, w1);
            }
        }
    }
    cplx *out = c.out + s * c.out_sstride;
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
        const int m = o + n * N2;
        if (m >= c.M)
            break;
        const cplx ch = __ldg(&c.tab_out[m]);
        cz2_epilogue(c, s, m, cmul(acc[0][n], ch), cmul(acc[1][n], ch), out);
    }
}

template <class K>
static inline int cz2_launch(K kernel, const Cz2Args &a, unsigned grid, int nt, size_t smem, cudaStream_t st,
                             const char *name)
{
    {
        const int e = fnftb_smem_optin((const void *)kernel, smem);
        if (e != 0)
            return e;
    }
    if (g_fnftb_profile_on)
        fnftb_profile_begin(name, st);
    kernel<<<grid, nt, smem, st>>>(a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(st);
    ++g_fnftb_launch_count;
    return (int)cudaGetLastError();
}

template <class K>
static inline int cz2_launch_cluster(K kernel, const Cz2Args &a, unsigned items, unsigned CL, cudaStream_t st)
{
    const size_t smem = sizeof(cplx) << FNFTB_CZ2_ROW_L2;
    const int eo = fnftb_smem_optin((const void *)kernel, smem);
    if (eo != 0)
        return eo;
    cudaError_t e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(items * CL, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (g_fnftb_profile_on)
        fnftb_profile_begin("tree_up_last_cz", st);
    e = cudaLaunchKernelEx(&cfg, kernel, a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(st);
    ++g_fnftb_launch_count;
    return (int)e;
}

// Same contract as cz_run (chirpz_driver.cuh); a.vhat doubles as the permuted FFT(v).
int cz2_run(CzArgs c, cplx *tables, const TwSet &tw, cudaStream_t st, const Cz2SymSrc *src)
{
    static const int knob_pow2 = [] {  // FNFT_B200_CZ2_POW2=1: power-of-two lengths only (round-1 behaviour)
        const char *e = getenv("FNFT_B200_CZ2_POW2");
        return (e && e[0]) ? atoi(e) : 0;
    }();
    const size_t need = (size_t)c.deg + (size_t)c.M;
    int R = cz2_rows_for(need);
    if (knob_pow2)
        while (R & (R - 1))
            ++R;
    if (R == 0)
        return -1;
    const size_t L = (size_t)R << FNFTB_CZ2_ROW_L2;
    int l2L = 0;
    while (((size_t)1 << l2L) < L)
        ++l2L;
    c.L = (int)L;
    c.N2 = 1 << FNFTB_CZ2_ROW_L2;
    c.N1 = R;
    c.plan1 = make_fft_plan(1 << (l2L - FNFTB_CZ2_ROW_L2));  // only used by the table kernel's (unused) twiddle table
    c.tab_y = tables;
    c.tab_out = c.tab_y + (c.deg + 1);
    c.tab_ph = c.tab_out + c.M;
    c.tab_tw = c.tab_ph + 3 * (size_t)c.M;
    int rc;
    {
        long long tot = c.deg + 1;
        if (c.M > tot)
            tot = c.M;
        if ((long long)L > tot)
            tot = (long long)L;
        rc = launch_blocks<CzArgs, blk_cz_tables>(c, (unsigned)((tot + 255) / 256), 256, 0, st, "cz_filter");
        if (rc)
            return rc;
    }
    Cz2Args a;
    a.c = c;
    memset(&a.src, 0, sizeof(a.src));
    if (src)
        a.src = *src;
    a.tw = tw;
    a.vperm = c.vhat;
    a.l2L = l2L;
    a.R = R;
    a.L = (int)L;
    static const int knob_pf = [] {  // L2 prefetch distance of the three kernels (CTAs), 0 = off
        const char *e = getenv("FNFT_B200_PFD_CZ");
        return (e && e[0]) ? atoi(e) : 12;
    }();
    a.pf = knob_pf;
    const size_t smem = sizeof(cplx) << FNFTB_CZ2_ROW_L2;
    const unsigned cols1 = (unsigned)((size_t)1 << FNFTB_CZ2_ROW_L2) / 256;
    const size_t narr = (size_t)c.B * c.npoly;
#define CZ2_BY_R(KERNEL, ARGS, GRID, NAME)                                                     \
    switch (R) {                                                                               \
    case 2: rc = cz2_launch(KERNEL<2>, ARGS, GRID, 256, 0, st, NAME); break;                   \
    case 3: rc = cz2_launch(KERNEL<3>, ARGS, GRID, 256, 0, st, NAME); break;                   \
    case 4: rc = cz2_launch(KERNEL<4>, ARGS, GRID, 256, 0, st, NAME); break;                   \
    case 6: rc = cz2_launch(KERNEL<6>, ARGS, GRID, 256, 0, st, NAME); break;                   \
    case 8: rc = cz2_launch(KERNEL<8>, ARGS, GRID, 256, 0, st, NAME); break;                   \
    case 12: rc = cz2_launch(KERNEL<12>, ARGS, GRID, 256, 0, st, NAME); break;                 \
    default: rc = cz2_launch(KERNEL<16>, ARGS, GRID, 256, 0, st, NAME); break;                 \
    }
    // spectrum of the chirp filter (signal independent)
    {
        Cz2Args v = a;
        v.gen_v = 1;
        v.pf = 0;
        v.c.ybuf = c.ybuf;  // array 0 of the workspace as scratch
        CZ2_BY_R(k_cz2_cols_fwd, v, cols1, "cz_filter");
        if (rc)
            return rc;
        rc = cz2_launch(k_cz2_rows, v, (unsigned)R, 128, smem, st, "cz_filter");
        if (rc)
            return rc;
    }
    a.gen_v = 0;
    // 1: the whole last level as one cluster kernel.  OFF by default: measured 5.58 ms per 4096 signals against
    // 2.67 + 1.70 ms for k_up_rows_a + k_up_cols_cz -- with eight CTAs per cluster 7/8 of the column stage's reads are
    // remote shared-memory reads, which cost more than the 2 MB per signal of DRAM round trip they save (the cluster of
    // four of the N = 16384 level, k_up_smem_cluster<14>, does pay: 6.43 -> 5.52 ms)
    static const int knob_last_cluster = [] {
        const char *e = getenv("FNFT_B200_LAST_CLUSTER");
        return (e && e[0]) ? atoi(e) : 0;
    }();
    if (a.src.fused && a.src.up.rows_pending && knob_last_cluster) {
        const int RT = 1 << (a.src.up.l2n - a.src.up.l2row);
        const unsigned items = (unsigned)narr;
        switch (RT * 100 + R) {
        case 203: rc = cz2_launch_cluster(k_up_last_cluster_cz<2, 3>, a, items, 2, st); break;
        case 204: rc = cz2_launch_cluster(k_up_last_cluster_cz<2, 4>, a, items, 2, st); break;
        case 406: rc = cz2_launch_cluster(k_up_last_cluster_cz<4, 6>, a, items, 4, st); break;
        case 408: rc = cz2_launch_cluster(k_up_last_cluster_cz<4, 8>, a, items, 4, st); break;
        case 812: rc = cz2_launch_cluster(k_up_last_cluster_cz<8, 12>, a, items, 8, st); break;
        case 816: rc = cz2_launch_cluster(k_up_last_cluster_cz<8, 16>, a, items, 8, st); break;
        default: return -1066;
        }
    } else if (a.src.fused) {
        rc = up_rows_a_pending(a.src.up, st);
        if (rc)
            return rc;
        const int RT = 1 << (a.src.up.l2n - a.src.up.l2row);
        const unsigned grid = (unsigned)(narr * cols1);
        switch (RT * 100 + R) {
        case 203: rc = cz2_launch(k_up_cols_cz<2, 3>, a, grid, 256, 0, st, "tree_up_cols_cz"); break;
        case 204: rc = cz2_launch(k_up_cols_cz<2, 4>, a, grid, 256, 0, st, "tree_up_cols_cz"); break;
        case 406: rc = cz2_launch(k_up_cols_cz<4, 6>, a, grid, 256, 0, st, "tree_up_cols_cz"); break;
        case 408: rc = cz2_launch(k_up_cols_cz<4, 8>, a, grid, 256, 0, st, "tree_up_cols_cz"); break;
        case 812: rc = cz2_launch(k_up_cols_cz<8, 12>, a, grid, 256, 0, st, "tree_up_cols_cz"); break;
        case 816: rc = cz2_launch(k_up_cols_cz<8, 16>, a, grid, 256, 0, st, "tree_up_cols_cz"); break;
        default: return -1065;  // cz2_fused_cols_supported said yes for a pair that is not instantiated
        }
    } else {
        CZ2_BY_R(k_cz2_cols_fwd, a, (unsigned)(narr * cols1), "cz_cols_fwd");
    }
    if (rc)
        return rc;
    rc = cz2_launch(k_cz2_rows, a, (unsigned)(narr * (size_t)R), 128, smem, st, "cz_rows");
    if (rc)
        return rc;
    {
        // rows of the last column pass that hold wanted outputs (m < M)
        const int need_rows = (c.M + (1 << FNFTB_CZ2_ROW_L2) - 1) >> FNFTB_CZ2_ROW_L2;
        const unsigned grid = (unsigned)((size_t)c.B * cols1);
        static const int knob_prune = [] {
            const char *e = getenv("FNFT_B200_CZ2_PRUNE");
            return (e && e[0]) ? atoi(e) : 1;
        }();
        // output-pruned kernels: NOUT = number of computed rows, a divisor of R
        int nout = 0;
        if (knob_prune) {
            switch (R) {
            case 4: nout = need_rows <= 1 ? 1 : (need_rows <= 2 ? 2 : 0); break;
            case 6: nout = need_rows <= 3 ? 3 : 0; break;
            case 8: nout = need_rows <= 2 ? 2 : (need_rows <= 4 ? 4 : 0); break;
            case 12: nout = need_rows <= 4 ? 4 : (need_rows <= 6 ? 6 : 0); break;
            case 16: nout = need_rows <= 4 ? 4 : (need_rows <= 8 ? 8 : 0); break;
            default: break;
            }
        }
        switch (R * 100 + nout) {
        case 401: rc = cz2_launch(k_cz2_cols_inv_p<4, 1>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 402: rc = cz2_launch(k_cz2_cols_inv_p<4, 2>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 603: rc = cz2_launch(k_cz2_cols_inv_p<6, 3>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 802: rc = cz2_launch(k_cz2_cols_inv_p<8, 2>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 804: rc = cz2_launch(k_cz2_cols_inv_p<8, 4>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 1204: rc = cz2_launch(k_cz2_cols_inv_p<12, 4>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 1206: rc = cz2_launch(k_cz2_cols_inv_p<12, 6>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 1604: rc = cz2_launch(k_cz2_cols_inv_p<16, 4>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        case 1608: rc = cz2_launch(k_cz2_cols_inv_p<16, 8>, a, grid, 256, 0, st, "cz_cols_inv"); break;
        default: CZ2_BY_R(k_cz2_cols_inv, a, grid, "cz_cols_inv"); break;
        }
    }
#undef CZ2_BY_R
    return rc;
}
// General four-step chirp-z (any L = N1 * 4096, chirpz_driver.cuh) with the row stage of the fast
// path: k_cz2_rows does "row FFT * FFT(v) * inverse row FFT" about three times faster than the
// first-generation blk_cz_rows (register-fused stride-1 passes, derived twiddles).  Nothing else
// changes: the row kernel is self-contained as long as the filter spectrum comes from the same
// kernel (gen_v), which stores it in its own permuted order.
int cz_run_fastrows(CzArgs a, cplx *tables, const TwSet &tw, cudaStream_t st)
{
    const CzGeom g = cz_geometry(a.deg, a.M);
    if (g.N2 != (1 << FNFTB_CZ2_ROW_L2))
        return cz_run(a, tables, st);
    a.L = g.L;
    a.N1 = g.N1;
    a.N2 = g.N2;
    a.C = g.C;
    a.log2C = ilog2i((unsigned)g.C);
    a.plan1 = make_fft_plan(g.N1);
    a.plan2 = make_fft_plan(g.N2);
    const int nt = 256;
    int rc;
    a.tab_y = tables;
    a.tab_out = a.tab_y + (a.deg + 1);
    a.tab_ph = a.tab_out + a.M;
    a.tab_tw = a.tab_ph + 3 * (size_t)a.M;
    {
        long long tot = a.deg + 1;
        if (a.M > tot)
            tot = a.M;
        if (g.L > tot)
            tot = g.L;
        rc = launch_blocks<CzArgs, blk_cz_tables>(a, (unsigned)((tot + nt - 1) / nt), nt, 0, st, "cz_filter");
        if (rc)
            return rc;
    }
    Cz2Args r;
    memset(&r, 0, sizeof(r));
    r.tw = tw;
    r.vperm = a.vhat;
    r.l2L = ilog2i((unsigned)g.L);
    r.R = g.N1;
    r.L = g.L;
    const size_t smem = sizeof(cplx) << FNFTB_CZ2_ROW_L2;
    // 1. spectrum of the chirp filter: columns into array 0 of the workspace, rows -> vhat (permuted)
    {
        CzArgs v = a;
        v.gen_v = 1;
        v.fwd_only = 1;
        v.vhat = a.ybuf;
        rc = launch_blocks<CzArgs, blk_cz_cols_fwd, 256, 3>(v, (unsigned)(g.N2 / g.C), nt,
                                                            cz_cols_smem_bytes(g.C, g.N1, 1), st, "cz_filter");
        if (rc)
            return rc;
        r.c = a;
        r.gen_v = 1;
        rc = cz2_launch(k_cz2_rows, r, (unsigned)g.N1, 128, smem, st, "cz_filter");
        if (rc)
            return rc;
    }
    a.gen_v = 0;
    a.fwd_only = 0;
    // 2. forward columns of all polynomials
    rc = launch_blocks<CzArgs, blk_cz_cols_fwd, 256, 3>(a, (unsigned)((size_t)a.B * a.npoly * (g.N2 / g.C)), nt,
                                                        cz_cols_smem_bytes(g.C, g.N1, 1), st, "cz_cols_fwd");
    if (rc)
        return rc;
    // 3. rows
    r.c = a;
    r.gen_v = 0;
    r.pf = 6;  // L2 prefetch of the row six CTAs ahead
    rc = cz2_launch(k_cz2_rows, r, (unsigned)((size_t)a.B * a.npoly * g.N1), 128, smem, st, "cz_rows");
    if (rc)
        return rc;
    // 4. inverse columns + epilogue, on the transposed tile when the tile has a multiple of 32 columns
    // (FNFT_B200_CZ_INV_T=0: the column-major tile of the first generation)
    static const int knob_inv_t = [] {
        const char *e = getenv("FNFT_B200_CZ_INV_T");
        return (e && e[0]) ? atoi(e) : 1;
    }();
    int Ci = g.C;
    if (knob_inv_t) {
        while ((Ci * a.npoly) % 32 != 0 && 2 * Ci <= g.N2 && cz_cols_smem_bytes(2 * Ci, g.N1, a.npoly) <= (size_t)72 * 1024)
            Ci *= 2;
        if ((Ci * a.npoly) % 32 == 0) {
            a.C = Ci;
            a.log2C = ilog2i((unsigned)Ci);
            a.inv_t = 1;
        } else {
            Ci = g.C;
        }
    }
#ifndef FNFTB_CZ_INV_MINB
#define FNFTB_CZ_INV_MINB 3  // resident CTAs per SM the register budget of the inverse column kernel is cut for
#endif
    return launch_blocks<CzArgs, blk_cz_cols_inv, 256, FNFTB_CZ_INV_MINB>(a, (unsigned)((size_t)a.B * (g.N2 / Ci)), nt,
                                                                          cz_cols_smem_bytes(Ci, g.N1, a.npoly), st, "cz_cols_inv");
}
// the general four-step path with the first-generation row kernels (table twiddles): plain DFTs of the inverse
// transform, where 1e-15 matters more than speed
int cz_run_exact(CzArgs a, cplx *tables, cudaStream_t st) { return cz_run(a, tables, st); }
#else
int cz_run_exact(CzArgs a, cplx *tables, cudaStream_t st);
int cz2_run(CzArgs c, cplx *tables, const TwSet &tw, cudaStream_t st, const Cz2SymSrc *src = nullptr);
int cz_run_fastrows(CzArgs a, cplx *tables, const TwSet &tw, cudaStream_t st);
#endif
#endif  // !FNFTB_EMUL

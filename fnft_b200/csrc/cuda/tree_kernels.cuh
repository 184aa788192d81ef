// fnft_b200 -- block programs for the per-sample transfer-matrix construction and
// the 2x2 polynomial product tree.
//
// Replaces, for a whole batch of signals at once:
//   fnft__akns_fscatter   /root/reference/src/private/fnft__akns_fscatter.c:64-925
//                         (leaf cases :118-245, :402-433; zero-frequency expm :46-59)
//   fnft__poly_fmult2x2   /root/reference/src/private/fnft__poly_fmult.c:381-546
//                         (pair product :239-328, rescale :330-374)
//
// Device data layout ("level buffers"): level l holds, per signal, n_l matrices of
// degree d_l; element (signal s, matrix m, entry e in {11,12,21,22}, coefficient i)
// lives at  ((s*n_l + m)*4 + e)*(d_l+1) + i, coefficients highest power first
// exactly like the reference (fnft__poly_fmult.c:396-401).  Matrix m of level 0
// is sample D-1-m (fnft__akns_fscatter.c:408) so that the tree computes
// M_{D-1} * ... * M_0.
//
// Normalisation (fnft__poly_fmult.c:330-374,492-493) is applied lazily: the
// producer of a matrix records max|coeff| in mx[s*n_l+m]; the consumer scales
// its inputs by 2^-floor(log2(max)) on load and adds the exponents to W[s].
// Powers of two commute with everything, so result*2^W is unchanged.
#pragma once
#include "fft_core.cuh"

// akns scheme ids understood by the leaf kernel (values of
// fnft__akns_discretization_t, include/private/fnft__akns_discretization_t.h:43-72)
enum {
    FNFTB_AKNS_2SPLIT2_MODAL = 0,
    FNFTB_AKNS_2SPLIT1A = 1,
    FNFTB_AKNS_2SPLIT1B = 2,
    FNFTB_AKNS_2SPLIT2A = 3,
    FNFTB_AKNS_2SPLIT2B = 4,
    FNFTB_AKNS_2SPLIT2S = 5,
    FNFTB_AKNS_2SPLIT3S = 8,
    FNFTB_AKNS_2SPLIT4B = 10,
    FNFTB_AKNS_4SPLIT4B = 21
};

// how the second potential r is obtained
enum { FNFTB_R_NSE = 0 /* r = -kappa*conj(q) */, FNFTB_R_KDV = 1 /* r = -1 */,
       FNFTB_R_EXPLICIT = 2 };

// ---------------------------------------------------------------------------
// complex elementary functions needed by the leaves
// ---------------------------------------------------------------------------
HD cplx c_sqrt(cplx z)
{
    // principal square root, branch cut on the negative real axis
    const double ax = fabs(z.x), ay = fabs(z.y);
    if (ax == 0.0 && ay == 0.0)
        return make_cplx(0.0, z.y);
    const double m = hypot(z.x, z.y);
    double t = sqrt(0.5 * (m + ax));
    double re, im;
    if (z.x >= 0.0) {
        re = t;
        im = z.y / (2.0 * t);
    } else {
        re = ay / (2.0 * t);
        im = copysign(t, z.y);
    }
    return make_cplx(re, im);
}
HD cplx c_cos(cplx z)
{
    double s, c;
    SINCOS(z.x, &s, &c);
    return make_cplx(c * cosh(z.y), -s * sinh(z.y));
}
HD cplx c_sin(cplx z)
{
    double s, c;
    SINCOS(z.x, &s, &c);
    return make_cplx(s * cosh(z.y), c * sinh(z.y));
}
// sinc as in misc_CSINC (src/private/fnft__misc.c:306-314)
HD cplx c_sinc(cplx z)
{
    if (hypot(z.x, z.y) >= 1.0e-8)
        return cdiv(c_sin(z), z);
    return c_cos(cscale(z, 0.57735026918962576450914878050196 /* 1/sqrt(3) */));
}
// expm([0 q; r 0]*h): returns (cos D, q*h*sinc D, r*h*sinc D), D = h*sqrt(-q r)
// (fnft__akns_fscatter.c:46-59)
HD void zero_freq_expm(cplx *M, double h, cplx q, cplx r)
{
    const cplx Delta = cscale(c_sqrt(cneg(cmul(q, r))), h);
    const cplx del = cscale(c_sinc(Delta), h);
    M[0] = c_cos(Delta);
    M[1] = cmul(q, del);
    M[2] = cmul(r, del);
}

// ---------------------------------------------------------------------------
// Symmetric ("SYM") mode.  For the NSE (r = -kappa*conj(q)) every leaf, and hence every
// partial product, has the structure
//        M(z) = [  a(z)          b(z)   ]      f#(z) = z^d * conj(f(1/conj(z)))
//               [ -kappa*b#(z)   a#(z)  ]      (coefficients reversed and conjugated)
// so only the first row (a, b) is stored and multiplied: 4 forward + 2 inverse
// transforms per pair instead of 8 + 4, half the shared memory and half the HBM
// traffic.  On the unit circle at the N = 2d roots of unity  F#[k] = (-1)^k conj(F[k]).
// Level buffers then hold E = 2 entries (11, 12) per matrix.  Padding matrices are
// diag(z^d, 1) (they have the structure, z^d*I does not); the final kernel undoes the
// resulting shift of the second column and rebuilds the second row.
// ---------------------------------------------------------------------------

// ---------------------------------------------------------------------------
// leaf kernel: one thread per (signal, level-0 matrix)
// ---------------------------------------------------------------------------
struct LeafArgs {
    const cplx *q;   // [B][D]
    const cplx *r;   // [B][D] or NULL
    cplx *out;       // level-0 buffer
    double *mx;      // [B][npad]  (set to 1.0: leaves are not rescaled)
    int B, D, npad, deg0;
    int rmode, kappa, scheme;
    int sym;         // 1: store only the first row (entries 11, 12)
    double eps_t;
    int *status;     // per-signal status (nonzero = error), may be NULL
};

HD void leaf_matrix(cplx *p /*[4][deg0+1]*/, int scheme, int deg0, double eps_t, cplx q,
                    cplx r, int *err)
{
    cplx *p11 = p, *p12 = p + (deg0 + 1), *p21 = p + 2 * (deg0 + 1), *p22 = p + 3 * (deg0 + 1);
    const cplx Z = czero();
    switch (scheme) {
    case FNFTB_AKNS_2SPLIT2_MODAL: {  // fnft__akns_fscatter.c:118-147
        const double sclr = eps_t * hypot(q.x, q.y);
        if (q.x == r.x && sclr >= 1.0)
            *err = 1;
        const cplx one_m = csub(make_cplx(1.0, 0.0), cmul(cscale(q, eps_t), cscale(r, eps_t)));
        const cplx scl = cdiv(make_cplx(1.0, 0.0), c_sqrt(one_m));
        p11[0] = Z;
        p11[1] = scl;
        p12[0] = cmul(scl, cscale(q, eps_t));
        p12[1] = Z;
        p21[0] = Z;
        p21[1] = cmul(scl, cscale(r, eps_t));
        p22[0] = scl;
        p22[1] = Z;
        break;
    }
    case FNFTB_AKNS_2SPLIT1A: {  // :149-176
        cplx e[3];
        zero_freq_expm(e, eps_t / deg0, q, r);
        p11[0] = Z;
        p11[1] = e[0];
        p12[0] = Z;
        p12[1] = e[1];
        p21[0] = e[2];
        p21[1] = Z;
        p22[0] = e[0];
        p22[1] = Z;
        break;
    }
    case FNFTB_AKNS_2SPLIT1B:
    case FNFTB_AKNS_2SPLIT2A: {  // :178-203
        cplx e[3];
        zero_freq_expm(e, eps_t / deg0, q, r);
        p11[0] = Z;
        p11[1] = e[0];
        p12[0] = e[1];
        p12[1] = Z;
        p21[0] = Z;
        p21[1] = e[2];
        p22[0] = e[0];
        p22[1] = Z;
        break;
    }
    case FNFTB_AKNS_2SPLIT2B: {  // :204-230
        cplx e[3];
        zero_freq_expm(e, 0.5 * eps_t / deg0, q, r);
        p11[0] = cmul(e[1], e[2]);
        p11[1] = cmul(e[0], e[0]);
        p12[0] = cmul(e[0], e[1]);
        p12[1] = p12[0];
        p21[0] = cmul(e[0], e[2]);
        p21[1] = p21[0];
        p22[0] = p11[1];
        p22[1] = p11[0];
        break;
    }
    case FNFTB_AKNS_2SPLIT2S: {  // :232-258
        cplx e[3];
        zero_freq_expm(e, eps_t / deg0, q, r);
        p11[0] = Z;
        p11[1] = e[0];
        p12[0] = cscale(e[1], 0.5);
        p12[1] = p12[0];
        p21[0] = cscale(e[2], 0.5);
        p21[1] = p21[0];
        p22[0] = e[0];
        p22[1] = Z;
        break;
    }
    case FNFTB_AKNS_2SPLIT3S: {  // :331-360 (Burstein-Mirin; see leaf_chain.cuh for the chains)
        cplx a[3], b[3];
        zero_freq_expm(a, eps_t / deg0, q, r);
        zero_freq_expm(b, 2.0 * eps_t / deg0, q, r);
        const double third = 1.0 / 3.0, sixth = 1.0 / 6.0;
        p11[0] = cscale(cscale(cmul(a[1], a[2]), 2.0), third);
        p11[1] = Z;
        p11[2] = cscale(cadd(cscale(cmul(a[0], a[0]), 2.0), b[0]), third);
        p12[0] = cscale(csub(cscale(cmul(a[0], a[1]), 4.0), b[1]), sixth);
        p12[1] = cscale(cscale(b[1], 2.0), third);
        p12[2] = p12[0];
        p21[0] = cscale(csub(cscale(cmul(a[0], a[2]), 4.0), b[2]), sixth);
        p21[1] = cscale(cscale(b[2], 2.0), third);
        p21[2] = p21[0];
        p22[0] = p11[2];
        p22[1] = Z;
        p22[2] = p11[0];
        break;
    }
    case FNFTB_AKNS_2SPLIT4B:
    case FNFTB_AKNS_4SPLIT4B: {  // :402-433
        cplx a[3], b[3];
        zero_freq_expm(a, 0.5 * eps_t / deg0, q, r);
        zero_freq_expm(b, eps_t / deg0, q, r);
        const double third = 1.0 / 3.0;
        // products are written exactly in the reference's association order
        p11[0] = cscale(csub(cscale(cmul(cmul(b[0], a[1]), a[2]), 4.0), cmul(b[1], b[2])), third);
        p11[1] = cscale(cscale(cadd(cmul(cmul(b[1], a[0]), a[2]), cmul(cmul(b[2], a[0]), a[1])), 4.0), third);
        p11[2] = cscale(csub(cscale(cmul(cmul(b[0], a[0]), a[0]), 4.0), cmul(b[0], b[0])), third);
        p12[0] = cscale(csub(cscale(cmul(cmul(b[0], a[0]), a[1]), 4.0), cmul(b[0], b[1])), third);
        p12[1] = cscale(cscale(cadd(cmul(cmul(b[1], a[0]), a[0]), cmul(cmul(b[2], a[1]), a[1])), 4.0), third);
        p12[2] = p12[0];
        p21[0] = cscale(csub(cscale(cmul(cmul(b[0], a[0]), a[2]), 4.0), cmul(b[0], b[2])), third);
        p21[1] = cscale(cscale(cadd(cmul(cmul(b[2], a[0]), a[0]), cmul(cmul(b[1], a[2]), a[2])), 4.0), third);
        p21[2] = p21[0];
        p22[0] = p11[2];
        p22[1] = p11[1];
        p22[2] = p11[0];
        break;
    }
    default:
        *err = 2;
        for (int i = 0; i < 4 * (deg0 + 1); ++i)
            p[i] = Z;
        break;
    }
}

BLK void blk_leaf(const LeafArgs &a, blk3 bid, int nt, void * /*smem*/)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        const long long total = (long long)a.B * a.npad;
        if (gid < total) {
            const int s = (int)(gid / a.npad);
            const int m = (int)(gid % a.npad);
            const int d1 = a.deg0 + 1;
            cplx p[4 * 3];  // deg0 <= 2 for every scheme implemented here
            int err = 0;
            if (m < a.D) {
                const size_t idx = (size_t)s * a.D + (size_t)(a.D - 1 - m);
                const cplx q = a.q[idx];
                cplx r;
                if (a.rmode == FNFTB_R_NSE)
                    r = (a.kappa == 1) ? make_cplx(-q.x, q.y) : make_cplx(q.x, -q.y);
                else if (a.rmode == FNFTB_R_KDV)
                    r = make_cplx(-1.0, 0.0);
                else
                    r = a.r[idx];
                leaf_matrix(p, a.scheme, a.deg0, a.eps_t, q, r, &err);
            } else {
                // padding: z^deg * I (fnft__poly_fmult.c:422-438); diag(z^deg, 1) in SYM mode
                for (int i = 0; i < 4 * d1; ++i)
                    p[i] = czero();
                p[0] = make_cplx(1.0, 0.0);
                p[3 * d1] = make_cplx(1.0, 0.0);
            }
            const int E = a.sym ? 2 : 4;
            cplx *o = a.out + (size_t)gid * E * d1;
            for (int i = 0; i < E * d1; ++i)
                o[i] = p[i];
            a.mx[gid] = 1.0;
            if (err && a.status)
                a.status[s] = err;
        }
    }
}

// ---------------------------------------------------------------------------
// shared helpers for the pair-product kernels
// ---------------------------------------------------------------------------
// exponent a = floor(log2(max)) used by poly_rescale2x2 (fnft__poly_fmult.c:364)
HD int rescale_exponent(double mx)
{
    return (mx > 0.0 && mx < 1.0e308) ? floor_log2(mx) : 0;
}

#ifdef FNFTB_EMUL
static inline void atomic_max_double(double *p, double v)
{
    if (v > *p)
        *p = v;
}
static inline void atomic_add_int(int *p, int v) { *p += v; }
#else
DEV void atomic_max_double(double *p, double v)
{
    // valid for non-negative doubles (bit patterns are ordered like the values)
    atomicMax((unsigned long long *)p, (unsigned long long)__double_as_longlong(v));
}
DEV void atomic_add_int(int *p, int v) { atomicAdd(p, v); }
#endif

// ---------------------------------------------------------------------------
// Fused "last forward pass + pointwise 2x2 product + first inverse pass" for plans whose
// stride-1 pass has radix 4 (no twiddles at stride 1).  One call handles the 4 consecutive
// storage positions [pos0, pos0+4) of all operand arrays of ONE pair:
//   arr0   : operand array 0 of the pair; operand array p is at arr0 + p*astr
//   top    : top coefficient of operand array p at top[p*tstr]   (zero if no wrap)
//   sgn(j) : (-1)^k of the bin at position pos0+j = sgn_base for rows mode, else from
//            the first-pass digit (pos >> fs) & 1
// Results (E arrays) overwrite operand arrays 0..E-1 at the same positions, ready for the
// remaining inverse passes.
// ---------------------------------------------------------------------------
template <bool SYM>
HD void fused_pointwise4(cplx *arr0, size_t astr, int pos0, const cplx *top, int tstr, int wrap,
                         int fs, int rows_mode, double sgn_rows, int kappa)
{
    int ph[4];
    double sg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ph[j] = swz(pos0 + j);
        sg[j] = rows_mode ? sgn_rows : ((((pos0 + j) >> fs) & 1) ? -1.0 : 1.0);
    }
    if (SYM) {
        cplx v[4][4];  // [operand][position]
#pragma unroll
        for (int p = 0; p < 4; ++p) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[p][j] = arr0[p * astr + ph[j]];
            Dft<4, -1>::run(v[p]);
            if (wrap) {
                const cplx t = top[p * tstr];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[p][j].x += sg[j] * t.x;
                    v[p][j].y += sg[j] * t.y;
                }
            }
        }
        cplx c0[4], c1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // B21 = -kappa*(-1)^k conj(B12), B22 = (-1)^k conj(B11)
            const double ks = -(double)kappa * sg[j];
            c0[j] = cmul(v[0][j], v[2][j]);
            cfma(c0[j], v[1][j], make_cplx(ks * v[3][j].x, -ks * v[3][j].y));
            c1[j] = cmul(v[0][j], v[3][j]);
            cfma(c1[j], v[1][j], make_cplx(sg[j] * v[2][j].x, -sg[j] * v[2][j].y));
        }
        Dft<4, +1>::run(c0);
        Dft<4, +1>::run(c1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            arr0[ph[j]] = c0[j];
            arr0[astr + ph[j]] = c1[j];
        }
    } else {
        cplx A[4][4];  // A11, A12, A21, A22
#pragma unroll
        for (int p = 0; p < 4; ++p) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                A[p][j] = arr0[p * astr + ph[j]];
            Dft<4, -1>::run(A[p]);
            if (wrap) {
                const cplx t = top[p * tstr];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    A[p][j].x += sg[j] * t.x;
                    A[p][j].y += sg[j] * t.y;
                }
            }
        }
        cplx out[4][4];
#pragma unroll
        for (int col = 0; col < 2; ++col) {
            cplx B1[4], B2[4];  // B1c, B2c
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                B1[j] = arr0[(4 + col) * astr + ph[j]];
                B2[j] = arr0[(6 + col) * astr + ph[j]];
            }
            Dft<4, -1>::run(B1);
            Dft<4, -1>::run(B2);
            if (wrap) {
                const cplx t1 = top[(4 + col) * tstr], t2 = top[(6 + col) * tstr];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    B1[j].x += sg[j] * t1.x;
                    B1[j].y += sg[j] * t1.y;
                    B2[j].x += sg[j] * t2.x;
                    B2[j].y += sg[j] * t2.y;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                out[col][j] = cmul(A[0][j], B1[j]);       // C1c = A11*B1c + A12*B2c
                cfma(out[col][j], A[1][j], B2[j]);
                out[2 + col][j] = cmul(A[2][j], B1[j]);   // C2c = A21*B1c + A22*B2c
                cfma(out[2 + col][j], A[3][j], B2[j]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            Dft<4, +1>::run(out[e]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                arr0[e * astr + ph[j]] = out[e][j];
        }
    }
}

struct PairArgs {
    const cplx *in;       // level buffer, n_in matrices of degree d_in per signal
    cplx *out;            // level buffer, n_in/2 matrices of degree 2*d_in
    const double *mx_in;  // [B][n_in]
    double *mx_out;       // [B][n_in/2]  (must be zeroed when atomics are used)
    int *W;               // [B]
    cplx *gbuf;           // [B][pairs][E][R][N2] partial row results (R > 1)
    cplx *colbuf;         // [B][pairs][NA][R][N2] column-transformed operands (R > 1, use_col)
    int use_col;          // 1: blk_pair_cols ran before the rows kernel
    int B, n_in, d_in;
    int normalize;
    int kappa;    // SYM mode only
    int N;        // cyclic convolution length (power of two, >= 2*d_in)
    int wrap;     // 1 if N == 2*d_in (top coefficient handled analytically)
    int R, N2;    // N = R * N2; R == 1: whole product inside one CTA
    int G;        // pairs per CTA (R == 1), a power of two
    int log2N2, log2G;
    FftPlan plan; // plan for length N2
    TwTable T;
};

HD double load_scale(const PairArgs &a, size_t mat /* s*n_in + m */, int *expo)
{
    if (!a.normalize) {
        *expo = 0;
        return 1.0;
    }
    const int e = rescale_exponent(a.mx_in[mat]);
    *expo = e;
    return ldexp(1.0, -e);
}

// Small-degree pair products by direct convolution: one thread per
// (signal, pair, stored output entry).  d_in is a template parameter so that the
// accumulators stay in registers.
template <int DIN, bool SYM>
BLK void blk_pair_direct(const PairArgs &a, blk3 bid, int nt, void * /*smem*/)
{
    constexpr int E = SYM ? 2 : 4;
    FOR_THREADS(tid, nt)
    {
        const int npairs = a.n_in / 2;
        const long long gid = (long long)bid.x * nt + tid;
        const long long total = (long long)a.B * npairs * E;
        if (gid < total) {
            const int e = (int)(gid % E);
            const long long sp = gid / E;  // s*npairs + pair
            const int s = (int)(sp / npairs);
            const int row = SYM ? 0 : (e >> 1), col = SYM ? e : (e & 1);
            const size_t matA = (size_t)sp * 2;
            int eA, eB;
            const double sA = load_scale(a, matA, &eA);
            const double sB = load_scale(a, matA + 1, &eB);
            const cplx *A = a.in + matA * E * (DIN + 1);
            const cplx *Bm = A + E * (DIN + 1);
            const cplx *Ar0 = A + (SYM ? 0 : (row * 2 + 0)) * (DIN + 1);
            const cplx *Ar1 = A + (SYM ? 1 : (row * 2 + 1)) * (DIN + 1);
            cplx b0[DIN + 1], b1[DIN + 1];
            if (SYM) {
                // column c of B: (B11, B21) or (B12, B22) with B21 = -kappa*rc(B12),
                // B22 = rc(B11), rc = reverse + conjugate
                const cplx *B11 = Bm, *B12 = Bm + (DIN + 1);
#pragma unroll
                for (int j = 0; j <= DIN; ++j) {
                    if (col == 0) {
                        b0[j] = cscale(B11[j], sB);
                        const cplx t = B12[DIN - j];
                        b1[j] = cscale(make_cplx(t.x, -t.y), -(double)a.kappa * sB);
                    } else {
                        b0[j] = cscale(B12[j], sB);
                        const cplx t = B11[DIN - j];
                        b1[j] = cscale(make_cplx(t.x, -t.y), sB);
                    }
                }
            } else {
                const cplx *B0c = Bm + (0 * 2 + col) * (DIN + 1);
                const cplx *B1c = Bm + (1 * 2 + col) * (DIN + 1);
#pragma unroll
                for (int j = 0; j <= DIN; ++j) {
                    b0[j] = cscale(B0c[j], sB);
                    b1[j] = cscale(B1c[j], sB);
                }
            }
            cplx acc[2 * DIN + 1];
#pragma unroll
            for (int k = 0; k < 2 * DIN + 1; ++k)
                acc[k] = czero();
#pragma unroll
            for (int i = 0; i <= DIN; ++i) {
                const cplx a0 = cscale(Ar0[i], sA), a1 = cscale(Ar1[i], sA);
#pragma unroll
                for (int j = 0; j <= DIN; ++j) {
                    cfma(acc[i + j], a0, b0[j]);
                    cfma(acc[i + j], a1, b1[j]);
                }
            }
            cplx *o = a.out + ((size_t)sp * E + e) * (2 * DIN + 1);
            double m2 = 0.0;
#pragma unroll
            for (int k = 0; k < 2 * DIN + 1; ++k) {
                o[k] = acc[k];
                m2 = fmax(m2, cabs2(acc[k]));
            }
            atomic_max_double(&a.mx_out[sp], sqrt(m2));
            if (e == 0 && a.normalize)
                atomic_add_int(&a.W[s], eA + eB);
        }
    }
}

// FFT-based pair product.
//   R == 1 : CTA handles G consecutive pairs of one signal completely.
//            grid.x = B * ceil(npairs / G)
//   R  > 1 : CTA (s, pair, k1) computes the bins k == k1 (mod R) of the length-N
//            cyclic product from length-N2 transforms and writes the length-N2
//            inverse transform of those bins to gbuf; blk_pair_combine finishes.
//            grid.x = B * npairs * R
// NA = 8 (4 in SYM mode) operand arrays and NO = 4 (2) result arrays per pair.
// Shared memory: cplx S[NA][G][N2], cplx top[NA][G], cplx bot[NA][G] (SYM: index-0
// coefficients), double sc[2][G], double red[nt].
HD size_t pair_smem_bytes(int G, int N2, int nt, int sym)
{
    const size_t NA = sym ? 4 : 8;
    return sizeof(cplx) * (NA * G * N2 + 2 * NA * G) + sizeof(double) * (2 * G + (nt > G ? nt : G));
}

template <int MAXR, bool SYM>
BLK void blk_pair_fft_t(const PairArgs &a, blk3 bid, int nt, void *smem)
{
    constexpr int E = SYM ? 2 : 4;    // stored entries per matrix
    constexpr int NA = 2 * E;         // operand arrays per pair
    constexpr int L2E = SYM ? 1 : 2;  // log2(E)
    const int npairs = a.n_in / 2;
    const int G = a.G, N2 = a.N2, R = a.R;
    const int din1 = a.d_in + 1;
    cplx *S = (cplx *)smem;
    cplx *top = S + (size_t)NA * G * N2;
    cplx *bot = top + NA * G;
    double *sc = (double *)(bot + NA * G);
    double *red = sc + 2 * G;

    int s, pair0, k1;
    if (R == 1) {
        const int cpb = (npairs + G - 1) / G;  // CTAs per signal
        s = bid.x / cpb;
        pair0 = (bid.x % cpb) * G;
        k1 = 0;
    } else {
        k1 = bid.x % R;
        const int sp = bid.x / R;
        pair0 = sp % npairs;
        s = sp / npairs;
    }
    const int nbody = a.wrap ? a.d_in : din1;  // coefficients that enter the FFT
    const int l2n = a.log2N2, l2g = a.log2G;

    // phase 0: per-matrix scale factors, exponent bookkeeping; R > 1: w_R^(n1*k1)
    FOR_THREADS(tid, nt)
    {
        for (int t2 = tid; t2 < 2 * G; t2 += nt) {
            const int g = t2 >> 1, side = t2 & 1;
            const int pair = pair0 + g;
            double scale = 0.0;
            if (pair < npairs) {
                int e;
                scale = load_scale(a, (size_t)s * a.n_in + 2 * (size_t)pair + side, &e);
                if (a.normalize && k1 == 0 && e != 0)
                    atomic_add_int(&a.W[s], e);
            }
            sc[side * G + g] = scale;
        }
        if (R > 1) {
            for (int n1 = tid; n1 < R / 2; n1 += nt) {
                const cplx w = cispi(-2.0 * (double)((n1 * k1) % R) / (double)R);
                red[2 * n1] = w.x;
                red[2 * n1 + 1] = w.y;
            }
        }
    }
    BLOCK_SYNC();

    // phase A: load (and for R > 1 fold the radix-R column step into the load)
    FOR_THREADS(tid, nt)
    {
        if (R == 1) {
            // one warp per operand array pg = p*G + g, lanes over the coefficients
            const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
            for (int pg = warp; pg < NA * G; pg += nwarps) {
                const int g = pg & (G - 1), p = pg >> l2g;
                const int pair = pair0 + g;
                cplx *dst = S + ((size_t)pg << l2n);
                if (pair < npairs) {
                    const size_t mat = (size_t)s * a.n_in + 2 * (size_t)pair + (p >> L2E);
                    const cplx *x = a.in + (mat * E + (p & (E - 1))) * din1;
                    const double scl = sc[(p >> L2E) * G + g];
                    for (int i = lane; i < N2; i += 32)
                        dst[swz(i)] = (i < nbody) ? cscale(x[i], scl) : czero();
                } else {
                    for (int i = lane; i < N2; i += 32)
                        dst[swz(i)] = czero();
                }
            }
        } else if (a.use_col) {
            // operands were column-transformed (and scaled) by blk_pair_cols: plain row load
            const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
            const cplx *cb = a.colbuf + ((((size_t)s * npairs + pair0) * NA) * (size_t)R << l2n);
            for (int p = warp; p < NA; p += nwarps) {
                const cplx *x = cb + (((size_t)p * R + k1) << l2n);
                cplx *dst = S + ((size_t)p << l2n);
                for (int i = lane; i < N2; i += 32)
                    dst[swz(i)] = x[i];
            }
        } else {
            // y[n2] = w_N^(n2*k1) * sum_{n1<R/2} x[n1*N2+n2] * w_R^(n1*k1)
            const int half = R / 2;
            const size_t matA = (size_t)s * a.n_in + 2 * (size_t)pair0;
            const cplx *xbase = a.in + matA * E * din1;
            for (int n2 = tid; n2 < N2; n2 += nt) {
                const cplx wn = cispi(-2.0 * (double)((n2 * k1) & (a.N - 1)) / (double)a.N);
                cplx acc[NA];
#pragma unroll
                for (int p = 0; p < NA; ++p)
                    acc[p] = czero();
                // NA independent loads per n1 (the operand polynomials of the pair)
#pragma unroll 2
                for (int n1 = 0; n1 < half; ++n1) {
                    const int i = (n1 << l2n) + n2;
                    if (i >= nbody)
                        break;
                    const cplx w = make_cplx(red[2 * n1], red[2 * n1 + 1]);
#pragma unroll
                    for (int p = 0; p < NA; ++p)
                        cfma(acc[p], xbase[(size_t)p * din1 + i], w);
                }
#pragma unroll
                for (int p = 0; p < NA; ++p)
                    S[((size_t)p << l2n) + swz(n2)] = cscale(cmul(acc[p], wn), sc[(p >> L2E) * G]);
            }
        }
        // top coefficients (only used when wrap) and, in SYM mode, the index-0 ones
        for (int pg = tid; pg < NA * G; pg += nt) {
            const int g = pg & (G - 1), p = pg >> l2g;
            const int pair = pair0 + g;
            cplx v = czero(), v0 = czero();
            if (a.wrap && pair < npairs) {
                const size_t mat = (size_t)s * a.n_in + 2 * (size_t)pair + (p >> L2E);
                const cplx *x = a.in + (mat * E + (p & (E - 1))) * din1;
                v = cscale(x[a.d_in], sc[(p >> L2E) * G + g]);
                v0 = cscale(x[0], sc[(p >> L2E) * G + g]);
            }
            top[pg] = v;
            bot[pg] = v0;
        }
    }
    BLOCK_SYNC();

    // phase B: NA*G forward transforms of length N2 (without the stride-1 pass if fused)
    const int fs = plan_first_stride_log2(a.plan);
    const int fuse = (a.plan.npass >= 1 && a.plan.radix[a.plan.npass - 1] == 4) ? 1 : 0;
    if (MAXR == 16 && fuse) {
        FNFTB_SMEM_FFT_CT(-1, S, NA * G, a.plan, nt, a.T, 1);
    } else {
        FNFTB_SMEM_FFT_FWD_SKIP(S, NA * G, a.plan, nt, a.T, MAXR, fuse);
    }
    BLOCK_SYNC();

    // phase C: pointwise 2x2 products; results overwrite the A-side arrays
    if (fuse) {
        FOR_THREADS(tid, nt)
        {
            const int l2q = l2n - 2;  // position groups per array (log2)
            const int total = G << l2q;
            for (int idx = tid; idx < total; idx += nt) {
                const int grp = idx & ((1 << l2q) - 1);
                const int g = idx >> l2q;
                fused_pointwise4<SYM>(S + ((size_t)g << l2n), (size_t)G << l2n, grp * 4, top + g, G,
                                      a.wrap, fs, R > 1, (k1 & 1) ? -1.0 : 1.0, a.kappa);
            }
        }
    } else {
    FOR_THREADS(tid, nt)
    {
        const int total = G << l2n;
        for (int idx = tid; idx < total; idx += nt) {
            const int pos = idx & (N2 - 1);
            const int g = idx >> l2n;
            double sgn = 0.0;   // (-1)^k of the true frequency index k of this bin
            double sgnk = 1.0;  // same, but also defined without wrap (SYM needs wrap)
            if (R == 1)
                sgnk = ((pos >> fs) & 1) ? -1.0 : 1.0;
            else
                sgnk = (k1 & 1) ? -1.0 : 1.0;
            if (a.wrap)
                sgn = sgnk;
            const int ph = swz(pos);
            cplx v[NA];
#pragma unroll
            for (int p = 0; p < NA; ++p) {
                const cplx t = top[p * G + g];
                const cplx x = S[((size_t)(p * G + g) << l2n) + ph];
                v[p] = make_cplx(x.x + sgn * t.x, x.y + sgn * t.y);
            }
            if (SYM) {
                const cplx b21 = cscale(cconj(v[3]), -(double)a.kappa * sgnk);
                const cplx b22 = cscale(cconj(v[2]), sgnk);
                cplx c11 = cmul(v[0], v[2]);
                cfma(c11, v[1], b21);
                cplx c12 = cmul(v[0], v[3]);
                cfma(c12, v[1], b22);
                S[((size_t)(0 * G + g) << l2n) + ph] = c11;
                S[((size_t)(1 * G + g) << l2n) + ph] = c12;
            } else {
                constexpr int b = SYM ? 0 : 4;
                cplx c11 = cmul(v[0], v[b + 0]);
                cfma(c11, v[1], v[b + 2]);
                cplx c12 = cmul(v[0], v[b + 1]);
                cfma(c12, v[1], v[b + 3]);
                cplx c21 = cmul(v[2], v[b + 0]);
                cfma(c21, v[3], v[b + 2]);
                cplx c22 = cmul(v[2], v[b + 1]);
                cfma(c22, v[3], v[b + 3]);
                S[((size_t)(0 * G + g) << l2n) + ph] = c11;
                S[((size_t)(1 * G + g) << l2n) + ph] = c12;
                S[((size_t)(2 * G + g) << l2n) + ph] = c21;
                S[((size_t)(3 * G + g) << l2n) + ph] = c22;
            }
        }
    }
    }
    BLOCK_SYNC();

    // phase D: E*G inverse transforms (the stride-1 pass is already done if fused)
    if (MAXR == 16 && fuse) {
        FNFTB_SMEM_FFT_CT(+1, S, E * G, a.plan, nt, a.T, 1);
    } else {
        FNFTB_SMEM_FFT_INV_SKIP(S, E * G, a.plan, nt, a.T, MAXR, fuse);
    }
    BLOCK_SYNC();

    // phase E: write out
    if (R > 1) {
        FOR_THREADS(tid, nt)
        {
            cplx *gb = a.gbuf + (((size_t)s * npairs + pair0) * E) * (size_t)R * N2;
            for (int idx = tid; idx < E * N2; idx += nt) {
                const int n2 = idx & (N2 - 1), e = idx >> l2n;
                gb[((size_t)(e * R + k1) << l2n) + n2] = S[((size_t)e << l2n) + swz(n2)];
            }
        }
        return;
    }
    const int dout1 = 2 * a.d_in + 1;
    const double invN = 1.0 / (double)a.N;
    // red[g] collects max |c|^2 of pair g (G <= nt entries of red are free here)
    FOR_THREADS(tid, nt)
    {
        for (int g = tid; g < G; g += nt)
            red[g] = 0.0;
    }
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        // one warp per (pair g, entry e): lanes run over the coefficients
        const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
        for (int ge = warp; ge < G * E; ge += nwarps) {
            const int g = ge >> L2E, e = ge & (E - 1);
            const int pair = pair0 + g;
            if (pair >= npairs)
                continue;
            cplx ct = czero();
            if (a.wrap) {
                if (SYM) {
                    const cplx tA11 = top[0 * G + g], tA12 = top[1 * G + g];
                    if (e == 0) {
                        // top of B21 = -kappa*conj(B12[0])
                        ct = cmul(tA11, top[2 * G + g]);
                        cfma(ct, tA12, cscale(cconj(bot[3 * G + g]), -(double)a.kappa));
                    } else {
                        // top of B22 = conj(B11[0])
                        ct = cmul(tA11, top[3 * G + g]);
                        cfma(ct, tA12, cconj(bot[2 * G + g]));
                    }
                } else {
                    constexpr int b = SYM ? 0 : 4;
                    const int row = e >> 1, col = e & 1;
                    ct = cmul(top[(row * 2 + 0) * G + g], top[(b + col) * G + g]);
                    cfma(ct, top[(row * 2 + 1) * G + g], top[(b + 2 + col) * G + g]);
                }
            }
            const cplx *Se = S + ((size_t)(e * G + g) << l2n);
            cplx *oe = a.out + (((size_t)s * npairs + pair) * E + e) * dout1;
            double m2 = 0.0;
            const int nfft_out = a.wrap ? a.N : dout1;  // coefficients taken from the transform
            for (int i = lane; i < nfft_out; i += 32) {
                cplx v = cscale(Se[swz(i)], invN);
                if (i == 0)
                    v = csub(v, ct);  // ct == 0 unless wrap
                oe[i] = v;
                m2 = fmax(m2, cabs2(v));
            }
            if (a.wrap && lane == 0) {
                oe[a.N] = ct;
                m2 = fmax(m2, cabs2(ct));
            }
            m2 = WARP_MAX(m2);
#ifndef FNFTB_EMUL
            if (lane == 0)
#endif
                atomic_max_double(&red[g], m2);
        }
    }
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        for (int g = tid; g < G; g += nt)
            if (pair0 + g < npairs)
                a.mx_out[(size_t)s * npairs + pair0 + g] = sqrt(red[g]);
    }
}

BLK void blk_pair_fft(const PairArgs &a, blk3 bid, int nt, void *smem)
{
    blk_pair_fft_t<16, false>(a, bid, nt, smem);
}
BLK void blk_pair_fft_r8(const PairArgs &a, blk3 bid, int nt, void *smem)
{
    blk_pair_fft_t<8, false>(a, bid, nt, smem);
}
BLK void blk_pair_fft_sym(const PairArgs &a, blk3 bid, int nt, void *smem)
{
    blk_pair_fft_t<16, true>(a, bid, nt, smem);
}
BLK void blk_pair_fft_sym_r8(const PairArgs &a, blk3 bid, int nt, void *smem)
{
    blk_pair_fft_t<8, true>(a, bid, nt, smem);
}

BLK void blk_pair_fft_sym_r4(const PairArgs &a, blk3 bid, int nt, void *smem)
{
    blk_pair_fft_t<4, true>(a, bid, nt, smem);
}

// Column step of the four-step transform for the row-split levels: for every operand
// array (signal, pair, p) and column n2 the R-point DFT over n1 of x[n1*N2 + n2] (upper
// half of the inputs is zero), times w_N^(n2*k1), times the pending power-of-two scale.
// One thread per (array, n2); loads and stores are coalesced across n2.
// grid.x * nt = B * npairs * NA * N2
template <int R, bool SYM>
BLK void blk_pair_cols(const PairArgs &a, blk3 bid, int nt, void *)
{
    constexpr int E = SYM ? 2 : 4;
    constexpr int NA = 2 * E;
    constexpr int L2E = SYM ? 1 : 2;
    FOR_THREADS(tid, nt)
    {
        const int npairs = a.n_in / 2;
        const int N2 = a.N2, l2n = a.log2N2;
        const long long gid = (long long)bid.x * nt + tid;
        const int n2 = (int)(gid & (N2 - 1));
        const long long arr = gid >> l2n;  // (s*npairs + pair)*NA + p
        if (arr < (long long)a.B * npairs * NA) {
            const int p = (int)(arr % NA);
            const long long sp = arr / NA;
            const size_t mat = (size_t)sp * 2 + (p >> L2E);  // s*n_in + 2*pair + side
            const int din1 = a.d_in + 1;
            const int nbody = a.wrap ? a.d_in : din1;
            const cplx *x = a.in + (mat * E + (p & (E - 1))) * din1;
            int ex;
            const double scl = load_scale(a, mat, &ex);
            cplx v[R];
#pragma unroll
            for (int n1 = 0; n1 < R; ++n1) {
                const int i = (n1 << l2n) + n2;
                v[n1] = (n1 < R / 2 && i < nbody) ? cscale(x[i], scl) : czero();
            }
            Dft<R, -1>::run(v);
            const cplx w1 = cispi(-2.0 * (double)n2 / (double)a.N);
            cplx w = make_cplx(1.0, 0.0);
            cplx *o = a.colbuf + (((size_t)arr * R) << l2n) + n2;
#pragma unroll
            for (int k1 = 0; k1 < R; ++k1) {
                o[(size_t)k1 << l2n] = cmul(v[k1], w);
                w = cmul(w, w1);
            }
        }
    }
}

// Finishes a row-split product: radix-R inverse column step, 1/N scaling, wrap
// correction, max|coeff|.  One thread per (signal, pair, entry, n2); a CTA covers nt
// consecutive n2 of ONE (signal, pair, entry) (nt divides N2), so the max is reduced
// in shared memory and published with a single atomic per CTA.
template <int R, bool SYM>
BLK void blk_pair_combine(const PairArgs &a, blk3 bid, int nt, void *smem)
{
    constexpr int E = SYM ? 2 : 4;
    double *red = (double *)smem;
    const int npairs = a.n_in / 2;
    const int N2 = a.N2;
    const long long gid0 = (long long)bid.x * nt;
    const long long spe = gid0 >> a.log2N2;  // (s*npairs + pair)*E + e
    const int e = (int)(spe % E);
    const long long sp = spe / E;
    const int s = (int)(sp / npairs);
    const int pair = (int)(sp % npairs);
    FOR_THREADS(tid, nt)
    {
        const int n2 = (int)((gid0 + tid) & (N2 - 1));
        const cplx *gb = a.gbuf + (((size_t)spe * R) << a.log2N2);
        cplx v[R];
        {
            // w_N^(-k1*n2) by recurrence from w_N^(-n2) (R <= 64 steps)
            const cplx w1 = cispi(2.0 * (double)n2 / (double)a.N);
            cplx w = make_cplx(1.0, 0.0);
#pragma unroll
            for (int k1 = 0; k1 < R; ++k1) {
                v[k1] = cmul(gb[((size_t)k1 << a.log2N2) + n2], w);
                w = cmul(w, w1);
            }
        }
        Dft<R, +1>::run(v);
        const int dout1 = 2 * a.d_in + 1;  // == N + 1 when wrap, <= N otherwise
        const int din1 = a.d_in + 1;
        cplx *o = a.out + (size_t)spe * dout1;
        const double invN = 1.0 / (double)a.N;
        double m2 = 0.0;
        cplx ct = czero();
        if (n2 == 0 && a.wrap) {
            // product of the top coefficients of row (e>>1) of A and column (e&1) of B
            const size_t matA = (size_t)s * a.n_in + 2 * (size_t)pair;
            int eA, eB;
            const double sA = load_scale(a, matA, &eA), sB = load_scale(a, matA + 1, &eB);
            const cplx *A = a.in + matA * E * din1;
            const cplx *Bm = A + E * din1;
            cplx a0, a1, b0, b1;
            if (SYM) {
                a0 = cscale(A[0 * din1 + a.d_in], sA);
                a1 = cscale(A[1 * din1 + a.d_in], sA);
                if (e == 0) {
                    b0 = cscale(Bm[0 * din1 + a.d_in], sB);
                    b1 = cscale(cconj(Bm[1 * din1 + 0]), -(double)a.kappa * sB);
                } else {
                    b0 = cscale(Bm[1 * din1 + a.d_in], sB);
                    b1 = cscale(cconj(Bm[0 * din1 + 0]), sB);
                }
            } else {
                const int row = e >> 1, col = e & 1;
                a0 = cscale(A[(row * 2 + 0) * din1 + a.d_in], sA);
                a1 = cscale(A[(row * 2 + 1) * din1 + a.d_in], sA);
                b0 = cscale(Bm[(0 * 2 + col) * din1 + a.d_in], sB);
                b1 = cscale(Bm[(1 * 2 + col) * din1 + a.d_in], sB);
            }
            ct = cmul(a0, b0);
            cfma(ct, a1, b1);
            o[a.N] = ct;
            m2 = cabs2(ct);
        }
#pragma unroll
        for (int n1 = 0; n1 < R; ++n1) {
            const int i = (n1 << a.log2N2) + n2;
            if (i >= dout1)
                break;
            cplx c = cscale(v[n1], invN);
            if (i == 0)
                c = csub(c, ct);  // ct == 0 unless wrap
            o[i] = c;
            m2 = fmax(m2, cabs2(c));
        }
        // one atomic per warp (all lanes of a warp belong to the same pair)
        m2 = WARP_MAX(m2);
#ifndef FNFTB_EMUL
        if ((tid & 31) == 0)
#endif
            atomic_max_double(&a.mx_out[sp], sqrt(m2));
    }
    (void)red;
}

// Final step: apply the pending scale of the single remaining matrix, strip the
// padding-induced coefficients and emit [B][4][deg_out+1] plus W.  In SYM mode the
// second row is rebuilt from the first: T21 = -kappa*T12#, T22 = T11#, and the second
// column of the padded product is taken from the trailing coefficients.
struct FinalArgs {
    const cplx *in;      // level buffer with 1 matrix per signal, degree d_full
    const double *mx_in; // [B]
    cplx *tm;            // [B][4][deg_out+1]
    int *W;              // [B]
    int B, d_full, deg_out, normalize;
    int sym, kappa;
};

BLK void blk_tree_final(const FinalArgs &a, blk3 bid, int nt, void * /*smem*/)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        const long long per = 4LL * (a.deg_out + 1);
        const long long total = (long long)a.B * per;
        if (gid < total) {
            const int s = (int)(gid / per);
            const int rem = (int)(gid % per);
            const int e = rem / (a.deg_out + 1), i = rem % (a.deg_out + 1);
            int ex = 0;
            double scale = 1.0;
            if (a.normalize) {
                ex = rescale_exponent(a.mx_in[s]);
                scale = ldexp(1.0, -ex);
            }
            cplx v;
            if (!a.sym) {
                v = a.in[((size_t)s * 4 + e) * (a.d_full + 1) + i];
            } else {
                const cplx *r11 = a.in + ((size_t)s * 2 + 0) * (a.d_full + 1);
                const cplx *r12 = r11 + (a.d_full + 1);
                const int shift = a.d_full - a.deg_out;  // padding shifts column 2
                if (e == 0) {
                    v = r11[i];
                } else if (e == 1) {
                    v = r12[shift + i];
                } else if (e == 2) {
                    const cplx t = r12[shift + a.deg_out - i];
                    v = cscale(make_cplx(t.x, -t.y), -(double)a.kappa);
                } else {
                    const cplx t = r11[a.deg_out - i];
                    v = make_cplx(t.x, -t.y);
                }
            }
            a.tm[gid] = cscale(v, scale);
            if (rem == 0 && a.normalize)
                a.W[s] += ex;  // the tree kernels of this signal have all finished
        }
    }
}

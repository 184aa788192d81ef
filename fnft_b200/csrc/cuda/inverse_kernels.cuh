// fnft_b200 -- kernels of the INVERSE nonlinear Fourier transform:
//   * fast inverse scattering, fnft__nse_finvscatter (src/private/fnft__nse_finvscatter.c:70-366):
//     the reference recurses down to single samples and multiplies 2x2 polynomial matrices with FFTs on
//     every level.  Here sub-problems of degree <= 512 are solved by ONE CTA per signal with plain layer
//     peeling in shared memory (k_inv_block: n sequential steps, each updating all coefficients in parallel),
//     the levels above run the reference's four steps with batched FFT products (inverse_api.cu).
//   * Darboux transforms that add the discrete spectrum (src/fnft_nsev_inverse.c:680-905): independent per
//     time sample, one thread per (signal, sample); eigenfunctions of a seed potential (:908-1010): one
//     thread per (signal, eigenvalue, direction).
//   * elementwise steps of building the transfer matrix from a continuous spectrum (:251-678) and of the
//     spectral factorisation (src/private/fnft__poly_specfact.c:25-147).
#pragma once
#include "common.cuh"

// a batch of 2x2 polynomial matrices in device memory: entry e (11, 12, 21, 22), coefficient i (highest
// power first, like the reference) of signal s at p[s * sstride + e * estride + i]
struct InvPoly {
    cplx *p;
    size_t sstride, estride;
};

// ------------------------------------------------------------------------------------------------
// layer peeling of a block of n <= FNFTB_INV_BLOCK samples
// ------------------------------------------------------------------------------------------------
#define FNFTB_INV_BLOCK 512

struct InvBlockArgs {
    InvPoly T;    // in:  degree n; only used modulo z^(n+1) like the reference's recursion
    InvPoly Ti;   // out: degree n, Ti(z) T(z) = z^n (p == nullptr: not wanted)
    cplx *q;      // out: q[s * q_sstride + j], j < n
    size_t q_sstride;
    int *status;  // [B], set to 1 where 1 + kappa |Q|^2 <= 0 (src/private/fnft__nse_finvscatter.c:172-176)
    int n, kappa, modal;
    double eps_t;
};

// The transfer matrix of the block is M_(n-1)(z) ... M_0(z) with M_j = c_j [[1, Q_j z], [-kappa conj(Q_j), z]]
// (2SPLIT2A: Q = tan(eps |q|) q/|q|, c = cos(eps |q|); 2SPLIT2_MODAL: Q = eps q, c = 1/sqrt(1 + kappa |Q|^2)).
// With L_j(z) = c_j [[z, -Q_j z], [kappa conj(Q_j), 1]], L_j M_j = z I.  Step k reads Q of sample n-1-k from
// the constant coefficients (Q = -kappa conj(R21(0) / R11(0)), :160-163), replaces the remainder R by L R / z
// (degree m -> m-1) and the accumulated inverse A by L A (degree k -> k+1).  Both live in shared memory,
// indexed by POWER of z, in two copies (read one, write the other: one barrier per step).
template <int NT>
__global__ void __launch_bounds__(NT) k_inv_block(const InvBlockArgs a)
{
    extern __shared__ double2 fnftb_smem_inv[];
    const int n = a.n;
    const int LEN = n + 2;
    cplx *buf0 = (cplx *)fnftb_smem_inv;      // [8][LEN]: R11 R12 R21 R22 A11 A12 A21 A22
    cplx *buf1 = buf0 + 8 * LEN;
    const int tid = threadIdx.x;
    const size_t s = blockIdx.x;
    const double kap = (double)a.kappa;
    for (int i = tid; i < 8 * LEN; i += NT) {
        buf0[i] = czero();
        buf1[i] = czero();
    }
    __syncthreads();
    for (int e = 0; e < 4; ++e) {
        const cplx *src = a.T.p + s * a.T.sstride + e * a.T.estride;
        for (int i = tid; i <= n; i += NT)
            buf0[e * LEN + (n - i)] = src[i];
    }
    if (tid == 0) {
        buf0[4 * LEN] = make_cplx(1.0, 0.0);  // A = I
        buf0[7 * LEN] = make_cplx(1.0, 0.0);
    }
    __syncthreads();
    cplx *cur = buf0, *nxt = buf1;
    int bad = 0;
    for (int k = 0; k < n; ++k) {
        const int m = n - k;
        const cplx r11_0 = cur[0], r21_0 = cur[2 * LEN];
        const cplx t = cdiv(r21_0, r11_0);
        const cplx Q = make_cplx(-kap * t.x, kap * t.y);  // -kappa conj(t)
        const double aq2 = cabs2(Q);
        const double den = 1.0 + kap * aq2;
        if (!(den > 0.0)) {
            bad = 1;
            break;
        }
        const double scl = 1.0 / sqrt(den);
        if (tid == 0) {
            cplx qv;
            if (a.modal) {
                qv = make_cplx(Q.x / a.eps_t, Q.y / a.eps_t);
            } else {  // ATAN(absQ) * CEXP(I * CARG(Q)) / eps_t, :178
                const double aq = sqrt(aq2);
                const double f = (aq > 0.0) ? atan(aq) / (aq * a.eps_t) : 1.0 / a.eps_t;
                qv = make_cplx(Q.x * f, Q.y * f);
            }
            a.q[s * a.q_sstride + (size_t)(n - 1 - k)] = qv;
        }
        const cplx kQc = make_cplx(kap * Q.x, -kap * Q.y);  // kappa conj(Q)
        // remainder: powers 0 .. m-1
        for (int p = tid; p < m; p += NT) {
            const cplx x11 = cur[p], x12 = cur[LEN + p], x21 = cur[2 * LEN + p], x22 = cur[3 * LEN + p];
            const cplx y11 = cur[p + 1], y12 = cur[LEN + p + 1], y21 = cur[2 * LEN + p + 1], y22 = cur[3 * LEN + p + 1];
            nxt[p] = cscale(csub(x11, cmul(Q, x21)), scl);
            nxt[LEN + p] = cscale(csub(x12, cmul(Q, x22)), scl);
            nxt[2 * LEN + p] = cscale(cadd(cmul(kQc, y11), y21), scl);
            nxt[3 * LEN + p] = cscale(cadd(cmul(kQc, y12), y22), scl);
        }
        // accumulated inverse: powers 0 .. k+1
        for (int p = tid; p <= k + 1; p += NT) {
            const cplx x11 = cur[4 * LEN + p], x12 = cur[5 * LEN + p], x21 = cur[6 * LEN + p], x22 = cur[7 * LEN + p];
            cplx u11 = czero(), u12 = czero();
            if (p > 0) {
                const cplx w11 = cur[4 * LEN + p - 1], w12 = cur[5 * LEN + p - 1];
                const cplx w21 = cur[6 * LEN + p - 1], w22 = cur[7 * LEN + p - 1];
                u11 = cscale(csub(w11, cmul(Q, w21)), scl);
                u12 = cscale(csub(w12, cmul(Q, w22)), scl);
            }
            nxt[4 * LEN + p] = u11;
            nxt[5 * LEN + p] = u12;
            nxt[6 * LEN + p] = cscale(cadd(cmul(kQc, x11), x21), scl);
            nxt[7 * LEN + p] = cscale(cadd(cmul(kQc, x12), x22), scl);
        }
        __syncthreads();
        cplx *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    if (bad) {  // uniform across the CTA: every thread read the same two values
        if (tid == 0 && a.status)
            a.status[s] = 1;
        return;
    }
    if (a.Ti.p) {
        for (int e = 0; e < 4; ++e) {
            cplx *dst = a.Ti.p + s * a.Ti.sstride + e * a.Ti.estride;
            for (int i = tid; i <= n; i += NT)
                dst[i] = cur[(4 + e) * LEN + (n - i)];
        }
    }
}

// operands of a pair product: lev0[((s*2 + m)*4 + e)*(d+1) + i] = (m == 0 ? A : B) entry e, coefficient i;
// A is a polynomial of degree d - zlead stored without its zlead leading zeros
struct InvGatherArgs {
    cplx *lev0;
    InvPoly A, B;
    int d, zlead;
    long long total;  // nsignals * 8 * (d+1)
};
__global__ void __launch_bounds__(256) k_inv_gather(const InvGatherArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.total)
        return;
    const int len = a.d + 1;
    const int i = (int)(g % len);
    const long long r = g / len;
    const int e = (int)(r & 3), m = (int)((r >> 2) & 1);
    const size_t s = (size_t)(r >> 3);
    cplx v;
    if (m == 0)
        v = (i < a.zlead) ? czero() : a.A.p[s * a.A.sstride + e * a.A.estride + (i - a.zlead)];
    else
        v = a.B.p[s * a.B.sstride + e * a.B.estride + i];
    a.lev0[g] = v;
}

// dst entry e, coefficient j (j < count) = res[(s*4 + e)*rlen + i0 + j]
struct InvScatterArgs {
    const cplx *res;
    InvPoly dst;
    int rlen, i0, count;
    long long total;  // nsignals * 4 * count
};
__global__ void __launch_bounds__(256) k_inv_scatter(const InvScatterArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.total)
        return;
    const int j = (int)(g % a.count);
    const long long r = g / a.count;
    const int e = (int)(r & 3);
    const size_t s = (size_t)(r >> 2);
    a.dst.p[s * a.dst.sstride + e * a.dst.estride + j] = a.res[((size_t)s * 4 + e) * a.rlen + a.i0 + j];
}

// ------------------------------------------------------------------------------------------------
// discrete spectrum: Darboux transforms
// ------------------------------------------------------------------------------------------------
HD cplx cexp_c(cplx z)
{
    double sn, cs;
    SINCOS(z.y, &sn, &cs);
    const double e = exp(z.x);
    return make_cplx(e * cs, e * sn);
}

struct InvCdtArgs {
    const cplx *bs;   // [B][K] bound states, sorted by descending imaginary part
    const cplx *nc;   // [B][K] norming constants (residues already converted)
    cplx *q;          // [B][D]
    int B, K, D, zc;  // zc: first sample with t >= 0 (src/fnft_nsev_inverse.c:727-733)
    double T0, eps_t;
};

// pure multi-soliton (no continuous spectrum), src/fnft_nsev_inverse.c:803-846: per time sample the
// recursion over rho_k = b_k exp(2 i lambda_k t) for t >= 0, and its mirror image for t < 0
template <int KMAX>
__global__ void __launch_bounds__(128) k_inv_cdt_pure(const InvCdtArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)a.B * a.D)
        return;
    const int n = (int)(g % a.D);
    const size_t s = (size_t)(g / a.D);
    const cplx *bs = a.bs + s * a.K, *nc = a.nc + s * a.K;
    const double t = a.T0 + a.eps_t * (double)n;
    const bool right = (n >= a.zc);
    cplx rhok[KMAX];
    for (int i = 0; i < a.K; ++i) {
        const cplx l = bs[i];
        // exp(+-2 i lambda t)
        const cplx ex = right ? cexp_c(make_cplx(-2.0 * l.y * t, 2.0 * l.x * t))
                              : cexp_c(make_cplx(2.0 * l.y * t, -2.0 * l.x * t));
        const cplx c = right ? nc[i] : cdiv(make_cplx(1.0, 0.0), nc[i]);
        rhok[i] = cmul(c, ex);
    }
    cplx qt = czero();
    for (int i = 0; i < a.K; ++i) {
        const cplx rho = rhok[i];
        const cplx rhoc = cconj(rho);
        const cplx li = bs[i];
        const double ar = hypot(rho.x, rho.y);
        const double fi = 2.0 * li.y / (1.0 + ar * ar);  // f = i * fi
        // qt += 2 * rhoc * f * i = -2 fi rhoc
        qt.x -= 2.0 * fi * rhoc.x;
        qt.y -= 2.0 * fi * rhoc.y;
        const cplx f = make_cplx(0.0, fi);
        const cplx lic = cconj(li);
        for (int j = i + 1; j < a.K; ++j) {
            const cplx lj = bs[j];
            cplx num = cmul(csub(lj, li), rhok[j]);
            cfma(num, csub(rhok[j], rho), f);
            cplx one = make_cplx(1.0, 0.0);
            cfma(one, rhoc, rhok[j]);
            const cplx den = csub(csub(lj, lic), cmul(one, f));
            rhok[j] = cdiv(num, den);
        }
    }
    a.q[s * (size_t)a.D + n] = right ? qt : cconj(qt);
}

struct InvEigArgs {
    const cplx *bs;  // [B][K]
    const cplx *q;   // [B][D] seed potential
    cplx *phi;       // [B][2][K][D]  (phi1: index (0*K + i)*D + n, phi2: (K + i)*D + n), like the reference
    cplx *psi;       // [B][2][K][D]
    int B, K, D;
    double T0, T1;
};

HD cplx csqrt_c(cplx z)
{
    const double r = hypot(z.x, z.y);
    if (r == 0.0)
        return czero();
    double re = sqrt(0.5 * (r + fabs(z.x)));
    double im = 0.5 * z.y / re;
    if (z.x < 0.0) {
        const double t = re;
        re = fabs(im);
        im = (z.y < 0.0) ? -t : t;
    }
    return make_cplx(re, im);
}

// cosh(k h) and sinh(k h) / k for complex k
HD void ch_sh(cplx k, double h, cplx *ch, cplx *sh)
{
    const cplx z = make_cplx(k.x * h, k.y * h);
    double sn, cs;
    SINCOS(z.y, &sn, &cs);
    const double c = cosh(z.x), sv = sinh(z.x);
    *ch = make_cplx(c * cs, sv * sn);
    const cplx s = make_cplx(sv * cs, c * sn);
    *sh = cdiv(s, k);
}

// eigenfunctions of the seed potential at the bound states (src/fnft_nsev_inverse.c:908-1010): split-step
// with half steps; thread = (signal, eigenvalue, direction)
__global__ void __launch_bounds__(64) k_inv_eigenfunctions(const InvEigArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)a.B * a.K * 2)
        return;
    const int dir = (int)(g & 1);
    const int i = (int)((g >> 1) % a.K);
    const size_t s = (size_t)((g >> 1) / a.K);
    const int D = a.D, K = a.K;
    const cplx l = a.bs[s * K + i];
    const cplx *q = a.q + s * (size_t)D;
    const double h = ((a.T1 - a.T0) / (double)(D - 1)) * 0.5;
    const cplx l2 = cmul(l, l);
    if (dir == 0) {
        cplx *p1 = a.phi + (s * 2 * K + i) * (size_t)D, *p2 = a.phi + (s * 2 * K + K + i) * (size_t)D;
        cplx f1 = cexp_c(make_cplx(l.y * a.T0, -l.x * a.T0));  // exp(-i l T0)
        cplx f2 = czero();
        p1[0] = f1;
        p2[0] = f2;
        cplx qn = q[0];
        cplx ks = make_cplx(-cabs2(qn) - l2.x, -l2.y);
        cplx ch, sh;
        ch_sh(csqrt_c(ks), h, &ch, &sh);
        cplx u1 = cmuli(cmul(l, sh));
        for (int n = 1; n < D; ++n) {
            if (ks.x != 0.0 || ks.y != 0.0) {
                const cplx g1 = cadd(cmul(csub(ch, u1), f1), cmul(cmul(qn, sh), f2));
                const cplx g2 = cadd(cmul(cneg(cmul(cconj(qn), sh)), f1), cmul(cadd(ch, u1), f2));
                f1 = g1;
                f2 = g2;
            }
            qn = q[n];
            ks = make_cplx(-cabs2(qn) - l2.x, -l2.y);
            ch_sh(csqrt_c(ks), h, &ch, &sh);
            u1 = cmuli(cmul(l, sh));
            if (ks.x != 0.0 || ks.y != 0.0) {
                const cplx g1 = cadd(cmul(csub(ch, u1), f1), cmul(cmul(qn, sh), f2));
                const cplx g2 = cadd(cmul(cneg(cmul(cconj(qn), sh)), f1), cmul(cadd(ch, u1), f2));
                f1 = g1;
                f2 = g2;
            }
            p1[n] = f1;
            p2[n] = f2;
        }
    } else {
        cplx *p1 = a.psi + (s * 2 * K + i) * (size_t)D, *p2 = a.psi + (s * 2 * K + K + i) * (size_t)D;
        cplx f1 = czero();
        cplx f2 = cexp_c(make_cplx(-l.y * a.T1, l.x * a.T1));  // exp(i l T1)
        p1[D - 1] = f1;
        p2[D - 1] = f2;
        cplx qn = q[D - 1];
        cplx ks = make_cplx(-cabs2(qn) - l2.x, -l2.y);
        cplx ch, sh;
        ch_sh(csqrt_c(ks), h, &ch, &sh);
        cplx u1 = cmuli(cmul(l, sh));
        for (int n = D - 1; n > 0; --n) {
            if (ks.x != 0.0 || ks.y != 0.0) {
                const cplx qs = cmul(qn, sh), qcs = cmul(cconj(qn), sh);
                // scl = (ch - u1)(ch + u1) + |q|^2 sh^2
                cplx scl = cmul(csub(ch, u1), cadd(ch, u1));
                cfma(scl, qcs, qs);
                const cplx g1 = cdiv(csub(cmul(cadd(ch, u1), f1), cmul(qs, f2)), scl);
                const cplx g2 = cdiv(cadd(cmul(qcs, f1), cmul(csub(ch, u1), f2)), scl);
                f1 = g1;
                f2 = g2;
            }
            qn = q[n - 1];
            ks = make_cplx(-cabs2(qn) - l2.x, -l2.y);
            ch_sh(csqrt_c(ks), h, &ch, &sh);
            u1 = cmuli(cmul(l, sh));
            if (ks.x != 0.0 || ks.y != 0.0) {
                const cplx qs = cmul(qn, sh), qcs = cmul(cconj(qn), sh);
                cplx scl = cmul(csub(ch, u1), cadd(ch, u1));
                cfma(scl, qcs, qs);
                const cplx g1 = cdiv(csub(cmul(cadd(ch, u1), f1), cmul(qs, f2)), scl);
                const cplx g2 = cdiv(cadd(cmul(qcs, f1), cmul(csub(ch, u1), f2)), scl);
                f1 = g1;
                f2 = g2;
            }
            p1[n - 1] = f1;
            p2[n - 1] = f2;
        }
    }
}

struct InvDarbouxArgs {
    const cplx *bs, *nc;  // [B][K]
    const cplx *phi, *psi;  // [B][2][K][D]
    cplx *q;              // [B][D] in: seed, out: seed + solitons
    int B, K, D;
};

// Darboux update of the seed potential, one thread per (signal, sample): src/fnft_nsev_inverse.c:868-893
template <int KMAX>
__global__ void __launch_bounds__(128) k_inv_darboux(const InvDarbouxArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)a.B * a.D)
        return;
    const int n = (int)(g % a.D);
    const size_t s = (size_t)(g / a.D);
    const int K = a.K, D = a.D;
    const cplx *bs = a.bs + s * K, *nc = a.nc + s * K;
    const cplx *phi = a.phi + s * 2 * K * (size_t)D, *psi = a.psi + s * 2 * K * (size_t)D;
    cplx S1[KMAX], S2[KMAX];
    cplx qn = a.q[s * (size_t)D + n];
    for (int i = 0; i < K; ++i) {
        cplx phi1 = phi[n + (size_t)i * D], phi2 = phi[n + (size_t)(K + i) * D];
        cplx psi1 = psi[n + (size_t)i * D], psi2 = psi[n + (size_t)(K + i) * D];
        const cplx li = bs[i];
        for (int j = 0; j < i; ++j) {
            const cplx d1 = csub(li, S1[j]), d2 = csub(li, cconj(S1[j]));
            cplx t1 = csub(cmul(d1, phi1), cmul(S2[j], phi2));
            phi2 = cadd(cmul(cconj(S2[j]), phi1), cmul(d2, phi2));
            phi1 = t1;
            t1 = csub(cmul(d1, psi1), cmul(S2[j], psi2));
            psi2 = cadd(cmul(cconj(S2[j]), psi1), cmul(d2, psi2));
            psi1 = t1;
        }
        const cplx beta = cdiv(csub(phi1, cmul(nc[i], psi1)), csub(phi2, cmul(nc[i], psi2)));
        const double ab = hypot(beta.x, beta.y);
        const double t = ab * ab;
        const cplx lic = cconj(li);
        S1[i] = make_cplx((t * li.x + lic.x) / (1.0 + t), (t * li.y + lic.y) / (1.0 + t));
        // S2 = 2 i Im(l) beta / (1 + t)
        const double f = 2.0 * li.y / (1.0 + t);
        S2[i] = make_cplx(-f * beta.y, f * beta.x);
        // qn -= 2 i S2
        qn.x += 2.0 * S2[i].y;
        qn.y -= 2.0 * S2[i].x;
    }
    a.q[s * (size_t)D + n] = qn;
}

// ------------------------------------------------------------------------------------------------
// elementwise helpers of the continuous-spectrum path
// ------------------------------------------------------------------------------------------------
// out[s][n-1-k] = in[s][k]  (descending-order input of fnftb__dft)
__global__ void __launch_bounds__(256) k_inv_reverse(const cplx *in, cplx *out, long long B, int n)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * n)
        return;
    const int k = (int)(g % n);
    out[(g - k) + (n - 1 - k)] = in[g];
}

// ---- spectral factorisation (src/private/fnft__poly_specfact.c:25-147), batched; every kernel writes the
// REVERSED array that the next DFT (fnftb__dft, coefficients in descending order) reads ----
struct SfArgs {
    const cplx *in;   // kernel input  [B][Ms] (or the polynomials, see k_sf_load)
    const cplx *in2;  // second input  [B][Ms]
    cplx *out;        // kernel output [B][Ms]
    cplx *out2;
    size_t in_sstride, out_sstride;  // k_sf_load / k_sf_store: strides of the polynomial arrays
    long long B;
    int Ms, deg, kappa;
    int *warn;        // [B]
};
// zero-padded polynomial, reversed: out[s][Ms-1-i] = poly[s][i] (i <= deg), 0 otherwise
__global__ void __launch_bounds__(256) k_sf_load(const SfArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * a.Ms)
        return;
    const int i = (int)(g % a.Ms);
    const long long s = g / a.Ms;
    a.out[s * a.Ms + (a.Ms - 1 - i)] = (i <= a.deg) ? a.in[s * a.in_sstride + i] : czero();
}
// x_l = log|P|, 0.5 log(1 + |P|^2) or 0.5 log(1 - |P|^2) (:76-108): out = x, out2 = x reversed
__global__ void __launch_bounds__(256) k_sf_log(const SfArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * a.Ms)
        return;
    const int i = (int)(g % a.Ms);
    const long long s = g / a.Ms;
    const cplx P = a.in[g];
    const double ab = hypot(P.x, P.y);
    const double tol = 1.4901161193847656e-08;  // sqrt(eps)
    cplx x;
    if (a.kappa == 0) {
        if (ab < tol)
            a.warn[s] = 1;
        x = make_cplx(log(ab), 0.0);
    } else if (a.kappa < 0) {
        x = make_cplx(0.5 * log(1.0 + ab * ab), 0.0);
    } else {
        const double a2 = ab * ab;
        if (a2 > 1.0 - tol)
            a.warn[s] = 1;
        const double v = 1.0 - a2;
        // CLOG of a negative number: log|v| + i pi
        x = (v >= 0.0) ? make_cplx(0.5 * log(v), 0.0) : make_cplx(0.5 * log(-v), 0.5 * 3.14159265358979323846);
    }
    a.out[g] = x;
    a.out2[s * a.Ms + (a.Ms - 1 - i)] = x;
}
// Hilbert-transform multipliers (:116-121, including the zeroed bin M/2-1), reversed output
__global__ void __launch_bounds__(256) k_sf_hilbert(const SfArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * a.Ms)
        return;
    const int i = (int)(g % a.Ms);
    const long long s = g / a.Ms;
    const int M = a.Ms;
    const cplx v = a.in[g];
    cplx r;
    if (i == 0 || i == M / 2 - 1)
        r = czero();
    else if (i < M / 2 - 1)
        r = make_cplx(v.y / M, -v.x / M);   // * (-i / M)
    else
        r = make_cplx(-v.y / M, v.x / M);   // * (+i / M)
    a.out[s * M + (M - 1 - i)] = r;
}
// exp(x - i y) / M (:130-131), reversed output
__global__ void __launch_bounds__(256) k_sf_exp(const SfArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * a.Ms)
        return;
    const int i = (int)(g % a.Ms);
    const long long s = g / a.Ms;
    const cplx x = a.in[g], y = a.in2[g];
    // x - i y = (x.x + y.y) + i (x.y - y.x)
    const cplx e = cexp_c(make_cplx(x.x + y.y, x.y - y.x));
    a.out[s * a.Ms + (a.Ms - 1 - i)] = make_cplx(e.x / a.Ms, e.y / a.Ms);
}
// result[i] = conj(buf[deg - i]) (:136-137)
__global__ void __launch_bounds__(256) k_sf_store(const SfArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * (a.deg + 1))
        return;
    const int i = (int)(g % (a.deg + 1));
    const long long s = g / (a.deg + 1);
    a.out[s * a.out_sstride + i] = cconj(a.in[s * a.Ms + (a.deg - i)]);
}

// ---- transfer matrix from samples of the continuous spectrum (src/fnft_nsev_inverse.c:251-678) ----
struct InvTmArgs {
    const cplx *cs;   // [B][M] contspec (boundary phase factors already removed by the host) / DFT output
    cplx *out;        // kernel dependent
    long long B;
    int M, D, deg, kappa;
    double scale;     // B_OF_TAU: 2 eps_t / degree1step
};
// reordering for the FFT (:289-292) and reversal for fnftb__dft in one step
__global__ void __launch_bounds__(256) k_inv_cs_reorder(const InvTmArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * a.M)
        return;
    const int i = (int)(g % a.M);
    const long long s = g / a.M;
    const int M = a.M;
    const int src = (i <= M / 2) ? i + (M / 2 - 1) : i - (M / 2 + 1);
    a.out[s * M + (M - 1 - i)] = a.cs[s * M + src];
}
// entries 12 and 21 from the FFT b of the spectrum (:347-359, :601-610); entries 11 / 22: A(z) = 1 when unit_a
__global__ void __launch_bounds__(256) k_inv_tm_from_b(const InvTmArgs a, int unit_a)
{
    const int len = a.deg + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * len)
        return;
    const int i = (int)(g % len);
    const long long s = g / len;
    const cplx *b = a.cs + s * a.M;
    cplx *T = a.out + s * 4 * (long long)len;
    const int M = a.M, deg = a.deg;
    const int i0 = (deg <= M - 1) ? 0 : deg - (M - 1);
    cplx t12 = czero(), t21 = czero();
    if (i >= i0) {
        const cplx b1 = b[M - 1 - deg + i];
        t12 = make_cplx(-a.kappa * b1.x / M, a.kappa * b1.y / M);  // -kappa conj(b / M)
        const cplx b2 = b[deg - i];
        t21 = make_cplx(b2.x / M, b2.y / M);
    }
    T[len + i] = t12;
    T[2 * len + i] = t21;
    if (unit_a) {
        T[i] = (i == deg) ? make_cplx(1.0, 0.0) : czero();
        T[3 * len + i] = (i == 0) ? make_cplx(1.0, 0.0) : czero();
    }
}
// entry 22 = reversed entry 11 (:617-618)
__global__ void __launch_bounds__(256) k_inv_tm_mirror_a(const InvTmArgs a)
{
    const int len = a.deg + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * len)
        return;
    const int i = (int)(g % len);
    cplx *T = a.out + (g / len) * 4 * (long long)len;
    T[3 * len + i] = T[a.deg - i];
}
// B_OF_TAU (:650-672), step 1: b coefficients into entry 21 at offset 1
__global__ void __launch_bounds__(256) k_inv_tm_btau_b(const InvTmArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * a.D)
        return;
    const int i = (int)(g % a.D);
    const long long s = g / a.D;
    const int len = a.deg + 1;
    cplx v = a.cs[s * a.M + i];
    double f = a.scale;
    if (i == 0 || i == a.D - 1)
        f *= 0.5;
    cplx *T = a.out + s * 4 * (long long)len;
    T[2 * len + 1 + i] = make_cplx(v.x * f, v.y * f);
}
// step 2 (after the spectral factorisation wrote entry 11 at offset 1): entries 12, 22 and the four zeros
__global__ void __launch_bounds__(256) k_inv_tm_btau_rest(const InvTmArgs a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.B * a.D)
        return;
    const int i = (int)(g % a.D);
    const long long s = g / a.D;
    const int len = a.deg + 1, D = a.D;
    cplx *T = a.out + s * 4 * (long long)len;
    const cplx bb = T[2 * len + 1 + (D - 1 - i)];
    T[len + i] = make_cplx(-a.kappa * bb.x, a.kappa * bb.y);
    T[3 * len + i] = T[1 + (D - 1 - i)];
    if (i == 0) {
        T[0] = czero();
        T[2 * len - 1] = czero();
        T[2 * len] = czero();
        T[4 * len - 1] = czero();
    }
}

// ---- iterative construction of A(z), B(z) from a reflection coefficient, defocusing case
// (src/fnft_nsev_inverse.c:375-510, Algorithm 1 of arXiv:1607.01305v2); one signal, M = D = deg ----
struct InvIterArgs {
    const cplx *in;
    const cplx *in2;
    cplx *out;
    double *sum;  // accumulated mean phase change
    int D, kappa;
};
// out[D-1-i] = q / sqrt(1 + kappa |q|^2) / D (:434-438), reversed for fnftb__dft
__global__ void __launch_bounds__(256) k_it_prep(const InvIterArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.D)
        return;
    const cplx q = a.in[i];
    const double ab = hypot(q.x, q.y);
    const double f = 1.0 / (sqrt(1.0 + a.kappa * ab * ab) * a.D);
    a.out[a.D - 1 - i] = make_cplx(q.x * f, q.y * f);
}
// out[i] = in[D-1-i]
__global__ void __launch_bounds__(256) k_it_flip(const InvIterArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.D)
        return;
    a.out[i] = a.in[a.D - 1 - i];
}
// phase update (:457-466): in = IFFT of the reversed a coefficients, in2 = contspec (phase factors removed),
// out = reordered contspec with the new phases; sum += |arg| / D
__global__ void __launch_bounds__(256) k_it_phase(const InvIterArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double mine = 0.0;
    if (i < a.D) {
        const cplx v = a.in[i];
        const double ph = atan2(v.y, v.x);
        mine = fabs(ph) / a.D;
        const int M = a.D;
        const int src = (i <= M / 2) ? i + (M / 2 - 1) : i - (M / 2 + 1);
        double sn, cs;
        SINCOS(ph, &sn, &cs);
        a.out[i] = cmul(a.in2[src], make_cplx(cs, sn));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(a.sum, mine);
}
// natural-order reordering without reversal: out[i] = cs[reorder(i)]
__global__ void __launch_bounds__(256) k_it_reorder(const InvIterArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.D)
        return;
    const int M = a.D;
    a.out[i] = a.in[(i <= M / 2) ? i + (M / 2 - 1) : i - (M / 2 + 1)];
}
// transfer matrix from a_coeffs (in) and b_coeffs (in2), :492-502
__global__ void __launch_bounds__(256) k_it_build(const InvIterArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.D)
        return;
    const int D = a.D, len = D + 1;
    cplx *T = a.out;
    T[1 + i] = a.in[i];
    const cplx bb = a.in2[D - 1 - i];
    T[len + i] = make_cplx(-a.kappa * bb.x, a.kappa * bb.y);
    T[2 * len + 1 + i] = a.in2[i];
    T[3 * len + i] = a.in[D - 1 - i];
    if (i == 0) {
        T[0] = czero();
        T[2 * len - 1] = czero();
        T[2 * len] = czero();
        T[4 * len - 1] = czero();
    }
}

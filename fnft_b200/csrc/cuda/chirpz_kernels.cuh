// fnft_b200 -- batched chirp-z evaluation of transfer-matrix polynomials and the
// continuous-spectrum epilogues.
//
// Replaces fnft__poly_chirpz  (/root/reference/src/private/fnft__poly_chirpz.c:33-105)
// and the loops that follow its two calls in
//   nsev_compute_contspec     (/root/reference/src/fnft_nsev.c:822-876)
//   tf2contspec_negxi         (/root/reference/src/fnft_kdvv.c:169-203).
//
// Bluestein:  y_n = p[deg-n] A^-n W^(n^2/2),  v_n = W^(-n^2/2) (wrapped),
//             g = IFFT_L( FFT_L(y) . FFT_L(v) ),  result_m = W^(m^2/2) g_m / L.
// L is a power of two >= deg + M, factored L = N1 * N2 (four-step FFT):
//   cz_cols_fwd : generate y (or v) on the fly, N1-point column FFTs, twiddle
//   cz_rows     : N2-point row FFT, * FFT(v), inverse row FFT      (shared memory)
//   cz_cols_inv : conj twiddle, inverse column FFTs, * W^(m^2/2)/L, epilogue
// The spectrum of v depends only on (deg, M, W) and is computed once per batch.
// Spectra are kept in the digit-reversed order the in-place transforms produce.
#pragma once
#include "fft_core.cuh"

enum { FNFTB_CZ_RAW = 0, FNFTB_CZ_NSEV = 1, FNFTB_CZ_KDVV = 2 };

struct CzArgs {
    // polynomials: coefficient i (highest power first) of poly j of signal s is
    // tm[s*tm_sstride + ent[j]*(deg+1) + i]
    const cplx *tm;
    size_t tm_sstride;
    int ent[2];
    int npoly;
    int deg;
    int B, M;
    int L, N1, N2;
    int C;           // columns per CTA in the column kernels (power of two)
    int log2C;
    FftPlan plan1, plan2;
    double lwr, lwi; // ln|W|, arg W
    int dft_n;       // > 0: plain DFT of this length, W = exp(+-2 pi i / dft_n) and A = 1; the chirp factors are then
                     // formed from k^2 mod 2 dft_n with sincospi (exact argument reduction) instead of arg W * k^2 / 2,
                     // whose rounding of 2 pi / n costs ~1e-12 at n ~ 4000 (inverse transform, resampling)
    double lar, lai; // ln|A|, arg A
    cplx *ybuf;      // [B][npoly][N1][N2]
    int inv_t;       // blk_cz_cols_inv: 1 = transposed shared-memory tile [row][column] (C * npoly a multiple of 32)
    cplx *vhat;      // [N1][N2]
    // signal-independent tables, filled once per call by blk_cz_tables:
    cplx *tab_y;     // [deg+1]  A^-n W^(n^2/2)
    cplx *tab_out;   // [M]      W^(m^2/2) / L
    cplx *tab_tw;    // [N1][N2] four-step twiddle w_L^(n2*k1(pos)), row index = storage position
    cplx *tab_ph;    // [3][M]   epilogue phases exp(i*xi*ph_rho), exp(i*xi*ph_a), exp(i*xi*ph_b)
                     //          (KDVV: [0] = exp(2i*xi*kdv_ph), [1] = exp(i*xi*kdv_sqrtz))
    TwTable T;
    int gen_v;       // cz_cols_fwd: 1 = generate the chirp filter v instead of y
    int fwd_only;    // cz_rows: 1 = forward transform only (used for vhat)
    // epilogue
    int mode;        // FNFTB_CZ_*
    int cstype;      // NSEV: 0 reflection coefficient, 1 a and b, 2 both
    cplx *out;
    size_t out_sstride;
    size_t out_jstride;  // RAW mode: distance between the outputs of poly 0 and 1 (0 => M)
    const int *W;    // per-signal normalisation exponent (NSEV a/b), may be NULL
    double xi0, eps_xi;
    double ph_rho, ph_a, ph_b;  // NSEV boundary phase factors
    double kdv_ph;              // KDVV: T1 + boundary_coeff*eps_t
    double kdv_sqrtz;           // KDVV 2SPLIT2A correction: eps_t/deg, else 0
    int *status;                // [B], set to 3 (division by zero) when H0 == 0
};

// exp(lr*h_r) * exp(i*(li1*h1 + li2*h2)) with the phase accumulated in
// double-double so that huge arguments (n^2/2 * arg W) lose no accuracy.
HD cplx chirp_factor(double mag_arg, double li1, double h1, double li2, double h2)
{
    const double p1 = li1 * h1, e1 = fma(li1, h1, -p1);
    const double p2 = li2 * h2, e2 = fma(li2, h2, -p2);
    const double s = p1 + p2;
    const double bb = s - p1;
    const double t = (p1 - (s - bb)) + (p2 - bb);
    const double err = t + e1 + e2;
    double sn, cs;
    SINCOS(s, &sn, &cs);
    cplx r = make_cplx(cs - sn * err, sn + cs * err);
    if (mag_arg != 0.0)
        r = cscale(r, exp(mag_arg));
    return r;
}

// W^(+-k^2/2) of a DFT of length n, W = exp(sign 2 pi i / n): exp(sign i pi (k^2 mod 2n) / n)
HD cplx chirp_dft(long long k, int n, double sign)
{
    const long long m = (k * k) % (2LL * n);
    return cispi(sign * (double)m / (double)n);
}

// (column, row) of element idx of a [C columns][N1 rows] tile that lives column-major in shared memory (columns
// N1 * 16 bytes apart: a multiple of the 128-byte bank period) and row-major in global memory.  With consecutive
// lanes on consecutive columns every lane of a warp hits the same banks (round 1 ncu: 62 % of the shared-memory
// wavefronts of blk_cz_cols_inv were conflicts).  Here a warp covers 4 columns x 8 rows: the 8 rows fill one 128-byte
// bank period and the 4 columns cost the 4 wavefronts that 512 bytes need anyway, while the global accesses are
// 64-byte segments (two full sectors).
HD void cz_tile_index(int idx, int log2C, int N1, int *c, int *row)
{
    if (log2C >= 2 && N1 >= 8) {
        const int hi = idx >> 5;
        *c = (idx & 3) | ((hi & ((1 << (log2C - 2)) - 1)) << 2);
        *row = ((idx >> 2) & 7) | ((hi >> (log2C - 2)) << 3);
    } else {
        *c = idx & ((1 << log2C) - 1);
        *row = idx >> log2C;
    }
}

// fills the signal-independent tables; grid covers max(deg+1, M, L) elements
BLK void blk_cz_tables(const CzArgs &a, blk3 bid, int nt, void *)
{
    FOR_THREADS(tid, nt)
    {
        const long long i = (long long)bid.x * nt + tid;
        if (i <= a.deg) {
            const double dn = (double)i;
            a.tab_y[i] = a.dft_n > 0 ? chirp_dft(i, a.dft_n, a.lwi < 0.0 ? -1.0 : 1.0)
                                     : chirp_factor(-a.lar * dn + a.lwr * (0.5 * dn * dn), a.lwi, 0.5 * dn * dn,
                                                    -a.lai, dn);
        }
        if (i < a.M) {
            const double dm = (double)i;
            a.tab_out[i] = cscale(a.dft_n > 0 ? chirp_dft(i, a.dft_n, a.lwi < 0.0 ? -1.0 : 1.0)
                                              : chirp_factor(a.lwr * (0.5 * dm * dm), a.lwi, 0.5 * dm * dm, 0.0, 0.0),
                                  1.0 / (double)a.L);
            double sn, cs;
            if (a.mode == FNFTB_CZ_NSEV) {
                const double xi = a.xi0 + a.eps_xi * dm;
                SINCOS(xi * a.ph_rho, &sn, &cs);
                a.tab_ph[i] = make_cplx(cs, sn);
                SINCOS(xi * a.ph_a, &sn, &cs);
                a.tab_ph[a.M + i] = make_cplx(cs, sn);
                SINCOS(xi * a.ph_b, &sn, &cs);
                a.tab_ph[2 * (size_t)a.M + i] = make_cplx(cs, sn);
            } else if (a.mode == FNFTB_CZ_KDVV) {
                const double xi = -a.xi0 - dm * a.eps_xi;
                SINCOS(2.0 * xi * a.kdv_ph, &sn, &cs);
                a.tab_ph[i] = make_cplx(cs, sn);
                SINCOS(xi * a.kdv_sqrtz, &sn, &cs);
                a.tab_ph[a.M + i] = make_cplx(cs, sn);
            }
        }
        if (i < a.L) {
            const int pos = (int)(i >> ilog2i((unsigned)a.N2)), n2 = (int)(i & (a.N2 - 1));
            const int k1 = plan_freq_of_pos(a.plan1, pos);
            a.tab_tw[i] = cispi(-2.0 * (double)(((long long)n2 * k1) & (long long)(a.L - 1)) / (double)a.L);
        }
    }
}

HD size_t cz_cols_smem_bytes(int C, int N1, int npoly) { return sizeof(cplx) * (size_t)C * N1 * npoly; }

// grid.x = B * npoly * (N2 / C)   (gen_v: B = npoly = 1)
BLK void blk_cz_cols_fwd(const CzArgs &a, blk3 bid, int nt, void *smem)
{
    cplx *S = (cplx *)smem;
    const int C = a.C, N1 = a.N1, N2 = a.N2;
    const int tiles = N2 / C;
    const int tile = bid.x % tiles;
    const int sj = bid.x / tiles;  // s*npoly + j
    const int j = sj % a.npoly, s = sj / a.npoly;
    const int n2_0 = tile * C;
    const int Np = a.deg + 1;
    // Short polynomial, long transform (nsep rings: 8193 coefficients, L = 2^19): only the first
    // nz = ceil(Np / N2) of the N1 rows are non-zero, so the N1-point column transforms are
    // evaluated directly, X[k1] = sum_{n1 < nz} y[n1] w_N1^(n1 k1), instead of through the
    // shared-memory FFT -- 3 terms per output at that size; the kernel becomes a pure stream of
    // writes.
    const int nz = (Np + N2 - 1) / N2;
    if (!a.gen_v && nz <= 4 && N1 >= 32 && nz * (N1 + C) <= C * N1) {
        cplx *Wt = S;            // [nz][N1]  w_N1^(n1 * k1(pos)), indexed by storage position
        cplx *Y = S + nz * N1;   // [nz][C]
        FOR_THREADS(tid, nt)
        {
            for (int idx = tid; idx < nz * N1; idx += nt) {
                const int n1 = idx / N1, pos = idx - n1 * N1;
                const int k1 = plan_freq_of_pos(a.plan1, pos);
                Wt[idx] = cispi(-2.0 * (double)((n1 * k1) & (N1 - 1)) / (double)N1);
            }
            for (int idx = tid; idx < nz * C; idx += nt) {
                const int n1 = idx >> a.log2C, c = idx & (C - 1);
                const long long n = (long long)n1 * N2 + n2_0 + c;
                cplx v = czero();
                if (n < Np) {
                    const cplx p = a.tm[(size_t)s * a.tm_sstride + (size_t)a.ent[j] * Np + (a.deg - n)];
                    v = cmul(p, a.tab_y[n]);
                }
                Y[idx] = v;
            }
        }
        BLOCK_SYNC();
        FOR_THREADS(tid, nt)
        {
            cplx *dst = a.ybuf + (size_t)sj * a.L;
            for (int idx = tid; idx < C * N1; idx += nt) {
                const int c = idx & (C - 1), pos = idx >> a.log2C;
                const int n2 = n2_0 + c;
                cplx x = Y[c];  // n1 = 0: w = 1
                for (int n1 = 1; n1 < nz; ++n1)
                    cfma(x, Y[n1 * C + c], Wt[n1 * N1 + pos]);
                dst[(size_t)pos * N2 + n2] = cmul(x, a.tab_tw[(size_t)pos * N2 + n2]);
            }
        }
        return;
    }
    FOR_THREADS(tid, nt)
    {
        for (int idx = tid; idx < C * N1; idx += nt) {
            int c, n1;
            cz_tile_index(idx, a.log2C, N1, &c, &n1);
            const long long n = (long long)n1 * N2 + n2_0 + c;
            cplx v = czero();
            if (!a.gen_v) {
                if (n < Np) {
                    const cplx p = a.tm[(size_t)s * a.tm_sstride + (size_t)a.ent[j] * Np + (a.deg - n)];
                    v = cmul(p, a.tab_y[n]);
                }
            } else {
                // fnft__poly_chirpz.c:76-82
                double dn = -1.0;
                if (n < a.M)
                    dn = (double)n;
                else if (n > (long long)a.L - Np)
                    dn = (double)(a.L - n);
                if (dn >= 0.0)
                    v = a.dft_n > 0 ? chirp_dft((long long)dn, a.dft_n, a.lwi < 0.0 ? 1.0 : -1.0)
                                    : chirp_factor(-a.lwr * (0.5 * dn * dn), -a.lwi, 0.5 * dn * dn, 0.0, 0.0);
            }
            S[(size_t)c * N1 + swz(n1)] = v;
        }
    }
    BLOCK_SYNC();
    FNFTB_SMEM_FFT_FWD(S, C, a.plan1, nt, a.T);
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        cplx *dst = (a.gen_v ? a.vhat : a.ybuf + (size_t)sj * a.L);
        for (int idx = tid; idx < C * N1; idx += nt) {
            int c, pos;
            cz_tile_index(idx, a.log2C, N1, &c, &pos);
            const int n2 = n2_0 + c;
            const cplx w = a.tab_tw[(size_t)pos * N2 + n2];
            dst[(size_t)pos * N2 + n2] = cmul(S[(size_t)c * N1 + swz(pos)], w);
        }
    }
}

// grid.x = B * npoly * N1  (fwd_only: N1)
BLK void blk_cz_rows(const CzArgs &a, blk3 bid, int nt, void *smem)
{
    cplx *S = (cplx *)smem;
    const int N2 = a.N2;
    cplx *row = (a.fwd_only ? a.vhat : a.ybuf) + (size_t)bid.x * N2;
    const int pos1 = bid.x % a.N1;
    FOR_THREADS(tid, nt)
    {
        for (int i = tid; i < N2; i += nt)
            S[swz(i)] = row[i];
    }
    BLOCK_SYNC();
    FNFTB_SMEM_FFT_FWD(S, 1, a.plan2, nt, a.T);
    BLOCK_SYNC();
    if (a.fwd_only) {
        FOR_THREADS(tid, nt)
        {
            for (int i = tid; i < N2; i += nt)
                row[i] = S[swz(i)];
        }
        return;
    }
    FOR_THREADS(tid, nt)
    {
        const cplx *vrow = a.vhat + (size_t)pos1 * N2;
        for (int i = tid; i < N2; i += nt)
            S[swz(i)] = cmul(S[swz(i)], vrow[i]);
    }
    BLOCK_SYNC();
    FNFTB_SMEM_FFT_INV(S, 1, a.plan2, nt, a.T);
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        for (int i = tid; i < N2; i += nt)
            row[i] = S[swz(i)];
    }
}

// One radix-R butterfly of an inverse pass (same arithmetic as fft_pass_butterfly<R, +1>) on a TRANSPOSED tile:
// element `pos` of column `col` lives at S[pos * NC + col].  The 32 lanes of a warp take the same butterfly of
// 32 consecutive columns, so every shared-memory access of a warp is 512 contiguous bytes whatever the stride,
// and the twiddle loads are uniform.  (ncu, round 2: the column-major tile of the first generation spent 49 % of
// its shared-memory wavefronts on bank conflicts in the stride-4 / stride-1 passes -- the kernel was bound by the
// shared-memory pipe at 96 %, 1.6 TB/s of DRAM.)
template <int R>
HD void cz_inv_butterfly_t(cplx *S, int NC, int col, int u, int log2s, const TwTable &T)
{
    const int s = 1 << log2s;
    const int g = u >> log2s;
    const int o = u & (s - 1);
    const int base = (g << (Log2R<R>::value + log2s)) + o;
    const int len = Log2R<R>::value + log2s;
    cplx v[R];
#pragma unroll
    for (int j = 0; j < R; ++j)
        v[j] = S[(size_t)(base + j * s) * NC + col];
    if (o != 0) {
#pragma unroll
        for (int j = 1; j < R; ++j)
            v[j] = cmul(v[j], tw_lookup<+1>(T, j * o, len));
    }
    Dft<R, +1>::run(v);
#pragma unroll
    for (int j = 0; j < R; ++j)
        S[(size_t)(base + j * s) * NC + col] = v[j];
}

// grid.x = B * (N2 / C); handles all npoly polynomials of a signal so that the
// epilogue can combine them.  Only output indices m < M are produced.
BLK void blk_cz_cols_inv(const CzArgs &a, blk3 bid, int nt, void *smem)
{
    cplx *S = (cplx *)smem;
    const int C = a.C, N1 = a.N1, N2 = a.N2;
    const int tiles = N2 / C;
    const int tile = bid.x % tiles;
    const int s = bid.x / tiles;
    const int n2_0 = tile * C;
    const int NC = C * a.npoly;  // columns of the tile
    const bool tr = (a.inv_t != 0);
    // element (column cc, row pos) of the tile
#define FNFTB_CZ_TILE(cc, pos) (tr ? S[(size_t)(pos) * NC + (cc)] : S[(size_t)(cc) * N1 + swz(pos)])
    FOR_THREADS(tid, nt)
    {
        for (int j = 0; j < a.npoly; ++j) {
            const cplx *src = a.ybuf + ((size_t)s * a.npoly + j) * a.L;
            for (int idx = tid; idx < C * N1; idx += nt) {
                int c, pos;
                if (tr) {  // lanes over consecutive columns: C * 16 contiguous bytes per row, global and shared
                    c = idx & (C - 1);
                    pos = idx >> a.log2C;
                } else {
                    cz_tile_index(idx, a.log2C, N1, &c, &pos);
                }
                const int n2 = n2_0 + c;
                const cplx w = cconj(a.tab_tw[(size_t)pos * N2 + n2]);
                FNFTB_CZ_TILE(j * C + c, pos) = cmul(src[(size_t)pos * N2 + n2], w);
            }
        }
    }
    BLOCK_SYNC();
    if (tr) {
        // the inverse passes of FNFTB_SMEM_FFT_INV (reverse plan order, strides 1, r_last, ...) on the transposed tile
        int l2s = 0;
        for (int p = a.plan1.npass - 1; p >= 0; --p) {
            const int R = a.plan1.radix[p];
            const int l2r = ilog2i(R);
            const int items = NC << (a.plan1.log2n - l2r);
            FOR_THREADS(tid, nt)
            {
                for (int idx = tid; idx < items; idx += nt) {
                    const int col = idx & (NC - 1), u = idx / NC;
                    switch (R) {
                    case 16: cz_inv_butterfly_t<16>(S, NC, col, u, l2s, a.T); break;
                    case 8: cz_inv_butterfly_t<8>(S, NC, col, u, l2s, a.T); break;
                    case 4: cz_inv_butterfly_t<4>(S, NC, col, u, l2s, a.T); break;
                    default: cz_inv_butterfly_t<2>(S, NC, col, u, l2s, a.T); break;
                    }
                }
            }
            BLOCK_SYNC();
            l2s += l2r;
        }
    } else {
        FNFTB_SMEM_FFT_INV(S, C * a.npoly, a.plan1, nt, a.T);
        BLOCK_SYNC();
    }
    FOR_THREADS(tid, nt)
    {
        for (int idx = tid; idx < C * N1; idx += nt) {
            int c, n1;
            if (tr) {
                c = idx & (C - 1);
                n1 = idx >> a.log2C;
            } else {
                cz_tile_index(idx, a.log2C, N1, &c, &n1);
            }
            const long long m = (long long)n1 * N2 + n2_0 + c;
            if (m >= a.M)
                continue;
            const cplx ch = a.tab_out[m];
            cplx H[2];
            H[1] = czero();
            for (int j = 0; j < a.npoly; ++j)
                H[j] = cmul(FNFTB_CZ_TILE(j * C + c, n1), ch);
            cplx *o = a.out + (size_t)s * a.out_sstride;
            if (a.mode == FNFTB_CZ_RAW) {
                const size_t js = a.out_jstride ? a.out_jstride : (size_t)a.M;
                for (int j = 0; j < a.npoly; ++j)
                    o[(size_t)j * js + m] = H[j];
            } else if (a.mode == FNFTB_CZ_NSEV) {
                // src/fnft_nsev.c:846-876; H[0] = H11 (a-poly), H[1] = H21 (b-poly); the
                // phases exp(i*xi*phi) come from tab_ph
                size_t off = 0;
                if (a.cstype == 0 || a.cstype == 2) {
                    if (H[0].x == 0.0 && H[0].y == 0.0) {
                        if (a.status)
                            a.status[s] = 3;
                        o[m] = make_cplx(NAN, NAN);
                    } else {
                        o[m] = cdiv(cmul(H[1], a.tab_ph[m]), H[0]);
                    }
                    off = a.M;
                }
                if (a.cstype == 1 || a.cstype == 2) {
                    const double scale = ldexp(1.0, a.W ? a.W[s] : 0);
                    o[off + m] = cmul(cscale(H[0], scale), a.tab_ph[a.M + m]);
                    o[off + a.M + m] = cmul(cscale(H[1], scale), a.tab_ph[2 * (size_t)a.M + m]);
                }
            } else {
                // src/fnft_kdvv.c:186-203; H[0] = H12, H[1] = H22, xi grid negated
                const double xi = -a.xi0 - (double)m * a.eps_xi;
                cplx h12 = H[0];
                if (a.kdv_sqrtz != 0.0)
                    h12 = cdiv(h12, a.tab_ph[a.M + m]);
                const cplx num = cmul(a.tab_ph[m], h12);
                // 2*i*xi*H22 - H12
                const cplx den = make_cplx(-2.0 * xi * H[1].y - h12.x, 2.0 * xi * H[1].x - h12.y);
                o[m] = cdiv(num, den);
            }
        }
    }
}
#undef FNFTB_CZ_TILE

// fnft_b200 -- fused kernel for the LOW part of the product tree: one CTA turns a block
// of S consecutive samples of one signal into their degree-(deg0*S) transfer matrix
// entirely in shared memory (leaf construction, direct-convolution levels, then
// FFT-convolution levels), so that these levels cause no HBM traffic beyond reading the
// samples and writing one matrix per block.
//
// Same mathematics as tree_kernels.cuh (leaf: fnft__akns_fscatter.c:116-433; pair
// product and rescaling: fnft__poly_fmult.c:239-374); see there for the SYM mode.
//
// Shared memory regions (cplx elements):
//   X[capX] : level data            capX = E*S*(deg0+1)
//   Y[capY] : FFT work arrays / level data of alternate direct levels
//                                    capY = max(capX, NA*S*deg0)
//   top[NA*pmax], bot[NA*pmax]      pmax = pairs at the first FFT level (<= S/2)
//   double mx[2][S] per-matrix max of the current / next level, double sc[S], int wsum
#pragma once
#include "tree_kernels.cuh"

struct LowArgs {
    const cplx *q;   // [B][D]
    const cplx *r;   // [B][D] or NULL
    cplx *out;       // [B][npad/S] matrices of degree deg0*S, E entries each
    double *mx_out;  // [B][npad/S]
    int *W;          // [B]
    int *status;     // [B]
    int B, D, npad, deg0;
    int rmode, kappa, scheme;
    int normalize;
    int S, log2S;    // samples per CTA (power of two, <= npad)
    double eps_t;
    TwTable T;
};

HD size_t low_capX(int E, int S, int deg0) { return (size_t)E * S * (deg0 + 1); }
HD size_t low_capY(int E, int S, int deg0)
{
    const size_t a = low_capX(E, S, deg0), b = (size_t)2 * E * S * deg0;
    return a > b ? a : b;
}
HD size_t low_smem_bytes(int sym, int S, int deg0)
{
    const int E = sym ? 2 : 4;
    const size_t NA = 2 * E;
    return sizeof(cplx) * (low_capX(E, S, deg0) + low_capY(E, S, deg0) + 2 * NA * (S / 2)) +
           sizeof(double) * (3 * (size_t)S) + 16;
}

// direct pair product on shared-memory level data: one thread per (pair, entry)
template <int DIN, bool SYM>
BLK void low_direct_level(const cplx *in, cplx *out, const double *mx_in, double *mx_out, int npairs,
                         int normalize, int kappa, int tid, int nt, int *wsum_local)
{
    constexpr int E = SYM ? 2 : 4;
    for (int idx = tid; idx < npairs * E; idx += nt) {
        const int e = idx % E, pair = idx / E;
        const int row = SYM ? 0 : (e >> 1), col = SYM ? e : (e & 1);
        int eA = 0, eB = 0;
        double sA = 1.0, sB = 1.0;
        if (normalize) {
            eA = rescale_exponent(mx_in[2 * pair]);
            eB = rescale_exponent(mx_in[2 * pair + 1]);
            sA = ldexp(1.0, -eA);
            sB = ldexp(1.0, -eB);
        }
        const cplx *A = in + (size_t)(2 * pair) * E * (DIN + 1);
        const cplx *Bm = A + E * (DIN + 1);
        const cplx *Ar0 = A + (SYM ? 0 : (row * 2 + 0)) * (DIN + 1);
        const cplx *Ar1 = A + (SYM ? 1 : (row * 2 + 1)) * (DIN + 1);
        cplx b0[DIN + 1], b1[DIN + 1];
        if (SYM) {
            const cplx *B11 = Bm, *B12 = Bm + (DIN + 1);
#pragma unroll
            for (int j = 0; j <= DIN; ++j) {
                if (col == 0) {
                    b0[j] = cscale(B11[j], sB);
                    b1[j] = cscale(cconj(B12[DIN - j]), -(double)kappa * sB);
                } else {
                    b0[j] = cscale(B12[j], sB);
                    b1[j] = cscale(cconj(B11[DIN - j]), sB);
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
        cplx *o = out + ((size_t)pair * E + e) * (2 * DIN + 1);
        double m2 = 0.0;
#pragma unroll
        for (int k = 0; k < 2 * DIN + 1; ++k) {
            o[k] = acc[k];
            m2 = fmax(m2, cabs2(acc[k]));
        }
        atomic_max_double(&mx_out[pair], sqrt(m2));
        if (e == 0)
            *wsum_local += eA + eB;
    }
}

// grid.x = B * (npad / S)
template <int MAXR, bool SYM>
BLK void blk_tree_low_t(const LowArgs &a, blk3 bid, int nt, void *smem)
{
    constexpr int E = SYM ? 2 : 4;
    constexpr int NA = 2 * E;
    constexpr int L2E = SYM ? 1 : 2;
    const int S = a.S, deg0 = a.deg0;
    const int nblk = a.npad >> a.log2S;
    const int s = bid.x / nblk, blk = bid.x % nblk;
    cplx *X = (cplx *)smem;
    cplx *Y = X + low_capX(E, S, deg0);
    cplx *top = Y + low_capY(E, S, deg0);
    cplx *bot = top + NA * (S / 2);
    double *mxA = (double *)(bot + NA * (S / 2));
    double *mxB = mxA + S;
    double *sc = mxB + S;
    int *wsum = (int *)(sc + S);

    // number of direct-convolution levels (d_in = deg0, 2*deg0, ... <= 8)
    int ndirect = 0;
    {
        int n = S, d = deg0;
        while (n >= 2 && d <= 8) {
            ++ndirect;
            n >>= 1;
            d <<= 1;
        }
    }
    // direct levels ping-pong between the two regions and must END in X
    cplx *cur = (ndirect & 1) ? Y : X;
    cplx *oth = (ndirect & 1) ? X : Y;
    double *mxc = mxA, *mxo = mxB;

    // ---- leaves ------------------------------------------------------------------
    FOR_THREADS(tid, nt)
    {
        if (tid == 0)
            *wsum = 0;
        const int d1 = deg0 + 1;
        for (int m = tid; m < S; m += nt) {
            const int mg = blk * S + m;  // level-0 matrix index within the signal
            cplx p[4 * 3];
            int err = 0;
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
                leaf_matrix(p, a.scheme, deg0, a.eps_t, q, r, &err);
            } else {
                for (int i = 0; i < 4 * d1; ++i)
                    p[i] = czero();
                p[0] = make_cplx(1.0, 0.0);
                p[3 * d1] = make_cplx(1.0, 0.0);
            }
            cplx *o = cur + (size_t)m * E * d1;
            for (int i = 0; i < E * d1; ++i)
                o[i] = p[i];
            mxc[m] = 1.0;
            if (err && a.status)
                a.status[s] = err;
        }
    }
    BLOCK_SYNC();

    int n = S, d = deg0;
    // ---- direct levels ---------------------------------------------------------------
    while (n >= 2 && d <= 8) {
        const int npairs = n / 2;
        FOR_THREADS(tid, nt)
        {
            for (int pidx = tid; pidx < npairs; pidx += nt)
                mxo[pidx] = 0.0;
        }
        BLOCK_SYNC();
        FOR_THREADS(tid, nt)
        {
            int wl = 0;
            switch (d) {
            case 1: low_direct_level<1, SYM>(cur, oth, mxc, mxo, npairs, a.normalize, a.kappa, tid, nt, &wl); break;
            case 2: low_direct_level<2, SYM>(cur, oth, mxc, mxo, npairs, a.normalize, a.kappa, tid, nt, &wl); break;
            case 4: low_direct_level<4, SYM>(cur, oth, mxc, mxo, npairs, a.normalize, a.kappa, tid, nt, &wl); break;
            default: low_direct_level<8, SYM>(cur, oth, mxc, mxo, npairs, a.normalize, a.kappa, tid, nt, &wl); break;
            }
            if (wl != 0)
                atomic_add_int(wsum, wl);
        }
        BLOCK_SYNC();
        cplx *t = cur;
        cur = oth;
        oth = t;
        double *tm = mxc;
        mxc = mxo;
        mxo = tm;
        n >>= 1;
        d <<= 1;
    }

    // ---- FFT levels: data in `cur` (== X when any FFT level follows), arrays in Y -----
    while (n >= 2) {
        const int npairs = n / 2;
        const int N = 2 * d;  // wrap level: d is a power of two
        const int l2n = ilog2i((unsigned)N);
        const int din1 = d + 1;
        const int dout1 = 2 * d + 1;
        FftPlan plan = make_fft_plan(N, MAXR);
        // scales + tops/bots
        FOR_THREADS(tid, nt)
        {
            int wl = 0;
            for (int m = tid; m < n; m += nt) {
                int e = 0;
                double scale = 1.0;
                if (a.normalize) {
                    e = rescale_exponent(mxc[m]);
                    scale = ldexp(1.0, -e);
                }
                sc[m] = scale;
                wl += e;
            }
            if (wl != 0)
                atomic_add_int(wsum, wl);
        }
        BLOCK_SYNC();
        FOR_THREADS(tid, nt)
        {
            // operand arrays: index pg = p*npairs + pair, p = side*E + entry; one warp per
            // array, lanes run over the coefficients (no divisions in the inner loop)
            const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
            const int l2p = ilog2i((unsigned)npairs);
            for (int pg = warp; pg < NA * npairs; pg += nwarps) {
                const int pair = pg & (npairs - 1), p = pg >> l2p;
                const int mat = 2 * pair + (p >> L2E);
                const cplx *x = cur + ((size_t)mat * E + (p & (E - 1))) * din1;
                const double scl = sc[mat];
                cplx *dst = Y + ((size_t)pg << l2n);
                for (int i = lane; i < N; i += 32)
                    dst[swz(i)] = (i < d) ? cscale(x[i], scl) : czero();
                if (lane == 0) {
                    top[pg] = cscale(x[d], scl);
                    bot[pg] = cscale(x[0], scl);
                }
            }
            for (int pidx = tid; pidx < npairs; pidx += nt)
                mxo[pidx] = 0.0;
        }
        BLOCK_SYNC();
        // forward transforms without the stride-1 pass, fused pointwise + first inverse
        // pass, remaining inverse passes (every FFT level here has N >= 32, so the plan
        // ends with the radix-4 pass)
        if constexpr (MAXR == 16) {
            FNFTB_SMEM_FFT_CT(-1, Y, NA * npairs, plan, nt, a.T, 1);
        } else {
            FNFTB_SMEM_FFT_FWD_SKIP(Y, NA * npairs, plan, nt, a.T, MAXR, 1);
        }
        BLOCK_SYNC();
        const int fs = plan_first_stride_log2(plan);
        FOR_THREADS(tid, nt)
        {
            const int l2q = l2n - 2;
            const int total = npairs << l2q;
            for (int idx = tid; idx < total; idx += nt) {
                const int grp = idx & ((1 << l2q) - 1);
                const int pair = idx >> l2q;
                fused_pointwise4<SYM>(Y + ((size_t)pair << l2n), (size_t)npairs << l2n, grp * 4,
                                      top + pair, npairs, 1, fs, 0, 1.0, a.kappa);
            }
        }
        BLOCK_SYNC();
        if constexpr (MAXR == 16) {
            FNFTB_SMEM_FFT_CT(+1, Y, E * npairs, plan, nt, a.T, 1);
        } else {
            FNFTB_SMEM_FFT_INV_SKIP(Y, E * npairs, plan, nt, a.T, MAXR, 1);
        }
        BLOCK_SYNC();
        // finalize into the (dead) input region: next level's data, n/2 matrices of degree
        // 2d.  One warp per (pair, entry): lanes run over the coefficients, the max is
        // reduced per warp in registers and published with one shared-memory atomic.
        FOR_THREADS(tid, nt)
        {
            const double invN = 1.0 / (double)N;
            const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
            for (int pe = warp; pe < npairs * E; pe += nwarps) {
                const int pair = pe >> L2E, e = pe & (E - 1);
                cplx ct;
                if (SYM) {
                    const cplx tA11 = top[0 * npairs + pair], tA12 = top[1 * npairs + pair];
                    if (e == 0) {
                        ct = cmul(tA11, top[2 * npairs + pair]);
                        cfma(ct, tA12, cscale(cconj(bot[3 * npairs + pair]), -(double)a.kappa));
                    } else {
                        ct = cmul(tA11, top[3 * npairs + pair]);
                        cfma(ct, tA12, cconj(bot[2 * npairs + pair]));
                    }
                } else {
                    constexpr int b = SYM ? 0 : 4;
                    const int row = e >> 1, col = e & 1;
                    ct = cmul(top[(row * 2 + 0) * npairs + pair], top[(b + col) * npairs + pair]);
                    cfma(ct, top[(row * 2 + 1) * npairs + pair], top[(b + 2 + col) * npairs + pair]);
                }
                const cplx *src = Y + ((size_t)(e * npairs + pair) << l2n);
                cplx *dst = cur + (size_t)pe * dout1;  // layout (pair*E + e)*dout1 + i
                double m2 = 0.0;
                for (int i = lane; i < N; i += 32) {
                    cplx v = cscale(src[swz(i)], invN);
                    if (i == 0)
                        v = csub(v, ct);
                    dst[i] = v;
                    m2 = fmax(m2, cabs2(v));
                }
                if (lane == 0) {
                    dst[N] = ct;
                    m2 = fmax(m2, cabs2(ct));
                }
                m2 = WARP_MAX(m2);
#ifndef FNFTB_EMUL
                if (lane == 0)
#endif
                    atomic_max_double(&mxo[pair], m2);  // squared; sqrt taken below
            }
        }
        BLOCK_SYNC();
        FOR_THREADS(tid, nt)
        {
            for (int pidx = tid; pidx < npairs; pidx += nt)
                mxo[pidx] = sqrt(mxo[pidx]);
        }
        BLOCK_SYNC();
        double *tm = mxc;
        mxc = mxo;
        mxo = tm;
        n >>= 1;
        d <<= 1;
    }

    // ---- write the block's matrix (degree d = deg0*S) ---------------------------------
    FOR_THREADS(tid, nt)
    {
        const size_t mo = (size_t)s * nblk + blk;
        cplx *o = a.out + mo * E * (d + 1);
        for (int idx = tid; idx < E * (d + 1); idx += nt)
            o[idx] = cur[idx];
        if (tid == 0) {
            a.mx_out[mo] = mxc[0];
            if (a.normalize && *wsum != 0)
                atomic_add_int(&a.W[s], *wsum);
        }
    }
}

BLK void blk_tree_low_sym(const LowArgs &a, blk3 bid, int nt, void *smem)
{
    blk_tree_low_t<16, true>(a, bid, nt, smem);
}
BLK void blk_tree_low_gen(const LowArgs &a, blk3 bid, int nt, void *smem)
{
    blk_tree_low_t<16, false>(a, bid, nt, smem);
}
BLK void blk_tree_low_sym_r8(const LowArgs &a, blk3 bid, int nt, void *smem)
{
    blk_tree_low_t<8, true>(a, bid, nt, smem);
}
BLK void blk_tree_low_gen_r8(const LowArgs &a, blk3 bid, int nt, void *smem)
{
    blk_tree_low_t<8, false>(a, bid, nt, smem);
}
BLK void blk_tree_low_sym_r4(const LowArgs &a, blk3 bid, int nt, void *smem)
{
    blk_tree_low_t<4, true>(a, bid, nt, smem);
}

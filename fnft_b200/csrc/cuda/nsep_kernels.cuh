// fnft_b200 -- kernels for the periodic NFT (fnft_nsep, GRIDSEARCH localization).
//
// Replaces, for a batch of signals:
//   de-rotation of the quasi-periodic phase    /root/reference/src/fnft_nsep.c:119-128
//   Floquet polynomials p_i = tm11_i + conj(tm11_{deg-i}), p_{deg/2} +- 2*2^-W
//                                               /root/reference/src/fnft_nsep.c:319-321,355
//   fnft__poly_roots_fftgridsearch 9-point scan + least-squares root estimate
//                                               /root/reference/src/private/fnft__poly_roots_fftgridsearch.c:80-147
//   z_to_lambda and the bounding-box filter     /root/reference/src/private/fnft__akns_discretization.c:225-240,
//                                               /root/reference/src/private/fnft__misc.c:114-157
// (the three chirp-z rings per polynomial come from chirpz_kernels.cuh).
#pragma once
#include "tree_kernels.cuh"

struct DerotArgs {
    const cplx *q;  // [B][D]
    cplx *out;      // [B][D]
    int B, D;
    double lam_shift, T0, eps_t;
};
BLK void blk_nsep_derotate(const DerotArgs &a, blk3 bid, int nt, void *)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        if (gid < (long long)a.B * a.D) {
            const int i = (int)(gid % a.D);
            double sn, cs;
            SINCOS(2.0 * a.lam_shift * (a.T0 + a.eps_t * i), &sn, &cs);
            a.out[gid] = cmul(a.q[gid], make_cplx(cs, sn));
        }
    }
}

struct FloquetArgs {
    const cplx *tm;  // [B][4][deg+1]
    const int *W;    // [B]
    cplx *P;         // [B][2][deg+1]: p_plus, p_minus
    int B, deg;
};
BLK void blk_nsep_polys(const FloquetArgs &a, blk3 bid, int nt, void *)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        const int d1 = a.deg + 1;
        if (gid < (long long)a.B * d1) {
            const int s = (int)(gid / d1), i = (int)(gid % d1);
            const cplx *t11 = a.tm + (size_t)s * 4 * d1;
            const cplx v = cadd(t11[i], cconj(t11[a.deg - i]));
            cplx pp = v, pm = v;
            if (i == a.deg / 2) {
                // p[deg/2] += 2*2^-W, then -= 4*2^-W (two roundings, as the reference)
                const double sh = 2.0 * ldexp(1.0, -a.W[s]);
                pp.x = v.x + sh;
                pm.x = pp.x - 4.0 * ldexp(1.0, -a.W[s]);
            }
            a.P[((size_t)s * 2 + 0) * d1 + i] = pp;
            a.P[((size_t)s * 2 + 1) * d1 + i] = pm;
        }
    }
}

HD cplx c_log(cplx z) { return make_cplx(log(hypot(z.x, z.y)), atan2(z.y, z.x)); }

struct ScanArgs {
    const cplx *vals;     // [B][npoly][3][M]  ring k=-1,0,+1
    int B, npoly, M;
    double PHI0, eps;     // grid: angle PHI0 + i*eps
    double lam_den;       // lambda = log(z) / (i*lam_den), lam_den = 2*eps_t/(deg1*up)
    int filtering;
    double box[4];
    double lam_shift;     // added to the kept values (src/fnft_nsep.c:196-203)
    int cap;              // capacity of out per (signal, poly)
    cplx *out;            // [B][npoly][cap] kept lambdas, scan order
    int *n_raw;           // [B][npoly] roots found before filtering
    int *n_kept;          // [B][npoly] after filtering (may exceed cap; only cap stored)
};

// evaluates the 9-point minimum-modulus test at grid index i and, if it passes,
// the least-squares linear root estimate.  Returns 1 if a root was produced.
HD int gridsearch_point(const ScanArgs &a, const cplx *v, int i, cplx *zr_out)
{
    const int M = a.M;
    const cplx y0 = v[M + i];
    const double tmp = hypot(y0.x, y0.y);
#define FNFTB_ABS(idx) hypot(v[(idx)].x, v[(idx)].y)
    if (tmp > FNFTB_ABS(i - 1)) return 0;
    if (tmp > FNFTB_ABS(i)) return 0;
    if (tmp > FNFTB_ABS(i + 1)) return 0;
    if (tmp > FNFTB_ABS(M + i - 1)) return 0;
    if (tmp > FNFTB_ABS(M + i + 1)) return 0;
    if (tmp > FNFTB_ABS(2 * M + i - 1)) return 0;
    if (tmp > FNFTB_ABS(2 * M + i)) return 0;
    if (tmp > FNFTB_ABS(2 * M + i + 1)) return 0;
#undef FNFTB_ABS
    double sn, cs;
    SINCOS(a.PHI0 + i * a.eps, &sn, &cs);
    const cplx z0 = make_cplx(cs, sn);
    cplx c = czero();
    double den = 0.0;
    for (int j = i - 1; j < i + 2; ++j) {
        for (int k = -1; k < 2; ++k) {
            if (j == 0 && k == 0)
                continue;  // sic: fnft__poly_roots_fftgridsearch.c:112-113
            SINCOS(a.PHI0 + j * a.eps, &sn, &cs);
            const double rad = 1.0 - k * a.eps;
            const cplx zi = make_cplx(rad * cs, rad * sn);
            const cplx yi = v[(k + 1) * M + j];
            const cplx dz = csub(zi, z0), dy = csub(yi, y0);
            // c += conj(dz)*dy
            c.x += dz.x * dy.x + dz.y * dy.y;
            c.y += dz.x * dy.y - dz.y * dy.x;
            const double ad = hypot(dz.x, dz.y);
            den += ad * ad;
        }
    }
    if (den == 0.0)
        return 0;
    c = make_cplx(c.x / den, c.y / den);
    cplx zr;
    if (c.x == 0.0 && c.y == 0.0) {
        if (y0.x != 0.0 || y0.y != 0.0)
            return 0;
        zr = z0;
    } else {
        zr = csub(z0, cdiv(y0, c));
        const cplx d = csub(zr, z0);
        if (hypot(d.x, d.y) > a.eps)
            return 0;
    }
    *zr_out = zr;
    return 1;
}

HD int scan_keep(const ScanArgs &a, cplx zr, cplx *lam_out)
{
    // lambda = log(z) / (i * lam_den)  ->  (arg z - i*ln|z|) / lam_den
    const cplx lg = c_log(zr);
    const cplx lam = make_cplx(lg.y / a.lam_den, -lg.x / a.lam_den);
    *lam_out = make_cplx(lam.x + a.lam_shift, lam.y);
    if (!a.filtering)
        return 1;
    return (lam.x >= a.box[0]) && (lam.x <= a.box[1]) && (lam.y >= a.box[2]) && (lam.y <= a.box[3]);
}

// One CTA per (signal, poly).  Each thread owns a contiguous range of grid indices;
// pass 1 counts, a prefix sum over the threads orders the output, pass 2 writes.
// shared memory: int cnt_raw[nt], cnt_kept[nt]
BLK void blk_nsep_scan(const ScanArgs &a, blk3 bid, int nt, void *smem)
{
    int *cnt_raw = (int *)smem;
    int *cnt_kept = cnt_raw + nt;
    const cplx *v = a.vals + (size_t)bid.x * 3 * a.M;
    const int npts = a.M - 2;  // indices 1 .. M-2
    const int per = (npts + nt - 1) / nt;
    FOR_THREADS(tid, nt)
    {
        int r = 0, k = 0;
        const int i0 = 1 + tid * per;
        int i1 = i0 + per;
        if (i1 > a.M - 1)
            i1 = a.M - 1;
        for (int i = i0; i < i1; ++i) {
            cplx zr, lam;
            if (gridsearch_point(a, v, i, &zr)) {
                ++r;
                k += scan_keep(a, zr, &lam);
            }
        }
        cnt_raw[tid] = r;
        cnt_kept[tid] = k;
    }
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        if (tid == 0) {
            int accr = 0, acck = 0;
            for (int t = 0; t < nt; ++t) {
                const int k = cnt_kept[t];
                accr += cnt_raw[t];
                cnt_kept[t] = acck;  // exclusive prefix
                acck += k;
            }
            a.n_raw[bid.x] = accr;
            a.n_kept[bid.x] = acck;
        }
    }
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        int off = cnt_kept[tid];
        const int i0 = 1 + tid * per;
        int i1 = i0 + per;
        if (i1 > a.M - 1)
            i1 = a.M - 1;
        cplx *o = a.out + (size_t)bid.x * a.cap;
        for (int i = i0; i < i1 && off < a.cap; ++i) {
            cplx zr, lam;
            if (gridsearch_point(a, v, i, &zr) && scan_keep(a, zr, &lam))
                o[off++] = lam;
        }
    }
}

#ifndef FNFTB_EMUL
// ---------------------------------------------------------------------------------------
// Round 2: the same scan with coalesced loads and one |p| per sample.  blk_nsep_scan gives every thread a
// contiguous range of grid indices (lanes 16 KB apart: every warp load touches 32 sectors) and evaluates
// hypot nine times per grid point, twice (count pass + write pass): 47 ms per 1024 signals of BASELINE config 5.
// Here a CTA walks over tiles of blockDim.x consecutive grid indices: |p| of the three rings goes to shared
// memory once per sample (coalesced 16-byte loads, next tile prefetched into registers), the 8 comparisons of
// the minimum-modulus test (fnft__poly_roots_fftgridsearch.c:86-104) read shared memory, and only the few
// points that pass run gridsearch_point (the same function as before: identical decisions and values).  The
// output keeps the scan order: warp ballots + per-warp counts give every kept value its slot.
// grid.x = B * npoly, blockDim.x = NT (multiple of 32), shared memory 3 * (NT + 2) doubles + 2 * NT / 32 ints
// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT, 2) k_nsep_scan_tiled(const ScanArgs a)
{
    __shared__ double habs[3][NT + 2];
    __shared__ int wraw[NT / 32], wkeep[NT / 32];
    const int M = a.M, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cplx *v = a.vals + (size_t)blockIdx.x * 3 * M;
    cplx *o = a.out + (size_t)blockIdx.x * a.cap;
    int base_raw = 0, base_keep = 0;
    // values of tile 0 (indices 1 + tid)
    cplx nx[3];
    {
        const int i = 1 + tid;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            nx[k] = (i < M) ? v[(size_t)k * M + i] : czero();
    }
    for (int t0 = 1; t0 < M - 1; t0 += NT) {
        const int i = t0 + tid;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            habs[k][tid + 1] = hypot(nx[k].x, nx[k].y);
        if (tid < 3) {  // left halo (index t0 - 1 >= 0) ...
            const cplx h = v[(size_t)tid * M + (t0 - 1)];
            habs[tid][0] = hypot(h.x, h.y);
        } else if (tid >= 32 && tid < 35) {  // ... and right halo (index t0 + NT), another warp
            const int k = tid - 32, j = t0 + NT;
            const cplx h = (j < M) ? v[(size_t)k * M + j] : czero();
            habs[k][NT + 1] = hypot(h.x, h.y);
        }
        {  // prefetch the next tile
            const int j = i + NT;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                nx[k] = (j < M) ? v[(size_t)k * M + j] : czero();
        }
        __syncthreads();
        int pass = 0;
        if (i < M - 1) {
            const double tmp = habs[1][tid + 1];
            pass = !(tmp > habs[0][tid]) && !(tmp > habs[0][tid + 1]) && !(tmp > habs[0][tid + 2]) &&
                   !(tmp > habs[1][tid]) && !(tmp > habs[1][tid + 2]) && !(tmp > habs[2][tid]) &&
                   !(tmp > habs[2][tid + 1]) && !(tmp > habs[2][tid + 2]);
        }
        int raw = 0, keep = 0;
        cplx lam = czero();
        if (pass) {
            cplx zr;
            if (gridsearch_point(a, v, i, &zr)) {
                raw = 1;
                keep = scan_keep(a, zr, &lam);
            }
        }
        const unsigned mraw = __ballot_sync(0xffffffffu, raw), mkeep = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) {
            wraw[warp] = __popc(mraw);
            wkeep[warp] = __popc(mkeep);
        }
        if (__syncthreads_or(raw)) {  // rare: a root in this tile
            int before = 0, tot_keep = 0, tot_raw = 0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) {
                before += (w < warp) ? wkeep[w] : 0;
                tot_keep += wkeep[w];
                tot_raw += wraw[w];
            }
            if (keep) {
                const int off = base_keep + before + __popc(mkeep & ((1u << lane) - 1u));
                if (off < a.cap)
                    o[off] = lam;
            }
            base_keep += tot_keep;
            base_raw += tot_raw;
            __syncthreads();  // wraw / wkeep are rewritten in the next tile
        }
    }
    if (tid == 0) {
        a.n_raw[blockIdx.x] = base_raw;
        a.n_kept[blockIdx.x] = base_keep;
    }
}
#endif

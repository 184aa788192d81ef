// fnft_b200 -- host-side driver of the product tree (level scheduling).
// Mirrors the level loop of fnft__poly_fmult2x2
// (/root/reference/src/private/fnft__poly_fmult.c:460-519) for a whole batch.
#pragma once
#include "launch.cuh"
#include "tree_kernels.cuh"

#ifdef FNFTB_EMUL
static inline void dev_memset0(void *p, size_t bytes, fnftb_stream_t) { memset(p, 0, bytes); }
#else
static inline void dev_memset0(void *p, size_t bytes, fnftb_stream_t st)
{
    cudaMemsetAsync(p, 0, bytes, st);
}
#endif

// Largest cyclic length whose 8 operand spectra fit one CTA's shared memory.
#ifndef FNFTB_TREE_SMEM_N
#define FNFTB_TREE_SMEM_N 1024
#endif

static inline size_t next_pow2_sz(size_t v)
{
    size_t r = 1;
    while (r < v)
        r *= 2;
    return r;
}

struct TreeWork {
    cplx *lev[2];    // level buffers, each >= B*4*npad*(deg0+1) cplx
    double *mx[2];   // per-matrix max|coeff|, each >= B*npad doubles
    cplx *gbuf;      // >= B*8*npad*deg0 cplx (row-split partial results)
    int *W;          // [B]
    int *status;     // [B]
};

// workspace sizes (elements) for B signals of npad matrices of degree deg0
static inline size_t tree_lev_elems(size_t B, size_t npad, size_t deg0)
{
    return B * 4 * npad * (deg0 + 1);
}
static inline size_t tree_gbuf_elems(size_t B, size_t npad, size_t deg0)
{
    return B * 8 * npad * deg0 + 64;
}

// import kernel: reference layout [4][n][deg+1] (entry-major, one signal) ->
// level-0 layout [n_pad][4][deg+1]; pads with z^deg * I.
struct ImportArgs {
    const cplx *p;
    cplx *out;
    double *mx;
    int n, npad, deg0;
};
BLK void blk_import(const ImportArgs &a, blk3 bid, int nt, void *)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        const int per = 4 * (a.deg0 + 1);
        if (gid < (long long)a.npad * per) {
            const int m = (int)(gid / per), rem = (int)(gid % per);
            const int e = rem / (a.deg0 + 1), i = rem % (a.deg0 + 1);
            cplx v;
            if (m < a.n)
                v = a.p[((size_t)e * a.n + m) * (a.deg0 + 1) + i];
            else
                v = make_cplx((i == 0 && (e == 0 || e == 3)) ? 1.0 : 0.0, 0.0);
            a.out[gid] = v;
            if (rem == 0)
                a.mx[m] = 1.0;
        }
    }
}

template <int DIN>
static inline int launch_direct(const PairArgs &pa, fnftb_stream_t st)
{
    const long long total = (long long)pa.B * (pa.n_in / 2) * 4;
    const int nt = 128;
    return launch_blocks<PairArgs, blk_pair_direct<DIN>, 128>(pa, (unsigned)((total + nt - 1) / nt), nt,
                                                         0, st, "tree_pair_direct");
}

template <int R>
static inline int launch_combine(const PairArgs &pa, fnftb_stream_t st)
{
    const long long total = (long long)pa.B * (pa.n_in / 2) * 4 * pa.N2;
    int nt = 128;
    if (nt > pa.N2)
        nt = pa.N2;  // a CTA must stay inside one (signal, pair, entry)
    return launch_blocks<PairArgs, blk_pair_combine<R>, 128>(pa, (unsigned)(total / nt), nt,
                                                        sizeof(double) * nt, st, "tree_pair_combine");
}

// Runs all levels.  On entry level buffer 0 holds npad matrices of degree deg0
// per signal (and mx[0] their max, normally 1.0).  On return *cur_out tells which
// level buffer holds the single product matrix (degree deg0*npad, pending scale in
// mx[*cur_out]).  use_direct=0 forces the FFT path on every level (testing).
static inline int tree_levels(const TreeWork &w, int B, int npad, int deg0, int normalize,
                              const TwTable &T, fnftb_stream_t st, int *cur_out,
                              int use_direct = 1, int smem_n = FNFTB_TREE_SMEM_N)
{
    int cur = 0;
    int n = npad, d = deg0;
    int rc = 0;
    while (n >= 2) {
        PairArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.in = w.lev[cur];
        pa.out = w.lev[1 - cur];
        pa.mx_in = w.mx[cur];
        pa.mx_out = w.mx[1 - cur];
        pa.W = w.W;
        pa.gbuf = w.gbuf;
        pa.B = B;
        pa.n_in = n;
        pa.d_in = d;
        pa.normalize = normalize;
        pa.T = T;
        const int npairs = n / 2;
        const bool direct =
            use_direct && (d == 1 || d == 2 || d == 3 || d == 4 || d == 6 || d == 8);
        if (direct) {
            dev_memset0(pa.mx_out, sizeof(double) * (size_t)B * npairs, st);
            switch (d) {
            case 1: rc = launch_direct<1>(pa, st); break;
            case 2: rc = launch_direct<2>(pa, st); break;
            case 3: rc = launch_direct<3>(pa, st); break;
            case 4: rc = launch_direct<4>(pa, st); break;
            case 6: rc = launch_direct<6>(pa, st); break;
            default: rc = launch_direct<8>(pa, st); break;
            }
        } else {
            int N = (int)next_pow2_sz((size_t)2 * d);
            pa.wrap = (N == 2 * d);
            if (!pa.wrap)
                N = (int)next_pow2_sz((size_t)2 * d + 1);
            pa.N = N;
            if (N <= smem_n) {
                pa.R = 1;
                pa.N2 = N;
                int G = smem_n / (2 * N);  // half of the big-CTA footprint per CTA
                if (G < 1)
                    G = 1;
                if (G > npairs)
                    G = npairs;
                pa.G = G;
                pa.log2G = ilog2i((unsigned)G);
                pa.log2N2 = ilog2i((unsigned)N);
                pa.plan = make_fft_plan(N);
                int nt = (8 * G * N) / 16;
                if (nt > 512)
                    nt = 512;
                if (nt < 64)
                    nt = 64;
                const unsigned grid = (unsigned)B * (unsigned)((npairs + G - 1) / G);
                rc = launch_blocks<PairArgs, blk_pair_fft, 512>(pa, grid, nt, pair_smem_bytes(G, N, nt),
                                                           st, "tree_pair_fft");
            } else {
                pa.R = N / smem_n;
                pa.N2 = smem_n;
                pa.G = 1;
                pa.log2G = 0;
                pa.log2N2 = ilog2i((unsigned)smem_n);
                pa.plan = make_fft_plan(smem_n);
                int nt = (8 * smem_n) / 16;
                if (nt > 512)
                    nt = 512;
                if (nt < 64)
                    nt = 64;
                const unsigned grid = (unsigned)B * (unsigned)npairs * (unsigned)pa.R;
                rc = launch_blocks<PairArgs, blk_pair_fft, 512>(pa, grid, nt,
                                                           pair_smem_bytes(1, smem_n, nt), st,
                                                           "tree_pair_fft_rows");
                if (rc)
                    return rc;
                dev_memset0(pa.mx_out, sizeof(double) * (size_t)B * npairs, st);
                switch (pa.R) {
                case 2: rc = launch_combine<2>(pa, st); break;
                case 4: rc = launch_combine<4>(pa, st); break;
                case 8: rc = launch_combine<8>(pa, st); break;
                case 16: rc = launch_combine<16>(pa, st); break;
                case 32: rc = launch_combine<32>(pa, st); break;
                case 64: rc = launch_combine<64>(pa, st); break;
                default: return -1000 - pa.R;  // transform too long for this build
                }
            }
        }
        if (rc)
            return rc;
        cur = 1 - cur;
        n /= 2;
        d *= 2;
    }
    *cur_out = cur;
    return 0;
}

static inline int tree_finalize(const TreeWork &w, int cur, int B, int d_full, int deg_out,
                                int normalize, cplx *tm, fnftb_stream_t st)
{
    FinalArgs fa;
    fa.in = w.lev[cur];
    fa.mx_in = w.mx[cur];
    fa.tm = tm;
    fa.W = w.W;
    fa.B = B;
    fa.d_full = d_full;
    fa.deg_out = deg_out;
    fa.normalize = normalize;
    const long long tot = (long long)B * 4 * (deg_out + 1);
    return launch_blocks<FinalArgs, blk_tree_final>(fa, (unsigned)((tot + 255) / 256), 256, 0, st,
                                                    "tree_final");
}

// Full fast scattering for a batch: leaves -> tree -> [B][4][deg_out+1] + W[B].
static inline int tree_fscatter(const TreeWork &w, const cplx *q, const cplx *r, int B, int D,
                                int deg0, int rmode, int kappa, int scheme, double eps_t,
                                int normalize, cplx *tm, const TwTable &T, fnftb_stream_t st,
                                int use_direct = 1, int smem_n = FNFTB_TREE_SMEM_N)
{
    const int npad = (int)next_pow2_sz((size_t)D);
    dev_memset0(w.W, sizeof(int) * (size_t)B, st);
    dev_memset0(w.status, sizeof(int) * (size_t)B, st);
    LeafArgs la;
    memset(&la, 0, sizeof(la));
    la.q = q;
    la.r = r;
    la.out = w.lev[0];
    la.mx = w.mx[0];
    la.B = B;
    la.D = D;
    la.npad = npad;
    la.deg0 = deg0;
    la.rmode = rmode;
    la.kappa = kappa;
    la.scheme = scheme;
    la.eps_t = eps_t;
    la.status = w.status;
    const long long total = (long long)B * npad;
    int rc = launch_blocks<LeafArgs, blk_leaf>(la, (unsigned)((total + 127) / 128), 128, 0, st,
                                               "tree_leaf");
    if (rc)
        return rc;
    int cur = 0;
    rc = tree_levels(w, B, npad, deg0, normalize, T, st, &cur, use_direct, smem_n);
    if (rc)
        return rc;
    return tree_finalize(w, cur, B, deg0 * npad, deg0 * D, normalize, tm, st);
}

// Product of n given matrices (one "signal"), reference layout in and out.
static inline int tree_fmult2x2(const TreeWork &w, const cplx *p_dev, int n, int deg0,
                                int normalize, cplx *tm, const TwTable &T, fnftb_stream_t st,
                                int use_direct = 1, int smem_n = FNFTB_TREE_SMEM_N)
{
    const int npad = (int)next_pow2_sz((size_t)n);
    dev_memset0(w.W, sizeof(int), st);
    ImportArgs ia;
    ia.p = p_dev;
    ia.out = w.lev[0];
    ia.mx = w.mx[0];
    ia.n = n;
    ia.npad = npad;
    ia.deg0 = deg0;
    const long long total = (long long)npad * 4 * (deg0 + 1);
    int rc = launch_blocks<ImportArgs, blk_import>(ia, (unsigned)((total + 127) / 128), 128, 0, st,
                                                   "tree_import");
    if (rc)
        return rc;
    int cur = 0;
    rc = tree_levels(w, 1, npad, deg0, normalize, T, st, &cur, use_direct, smem_n);
    if (rc)
        return rc;
    return tree_finalize(w, cur, 1, deg0 * npad, deg0 * n, normalize, tm, st);
}

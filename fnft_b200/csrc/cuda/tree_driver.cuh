// fnft_b200 -- host-side driver of the product tree (level scheduling).
// Mirrors the level loop of fnft__poly_fmult2x2
// (/root/reference/src/private/fnft__poly_fmult.c:460-519) for a whole batch.
#pragma once
#include "launch.cuh"
#include "tree_kernels.cuh"
#include "leaf_chain.cuh"
#include "tree_low_kernel.cuh"
#include "tree_low2.cuh"
#include "tree_up.cuh"
#include "tree_convert.cuh"

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
    cplx *colbuf;    // same size (column-transformed operands of the row-split levels)
    int *W;          // [B]
    int *status;     // [B]
    void *tt[2];     // spectrum path: top/bottom coefficients per matrix (Low2Tops), ping-pong
    const void *tws; // spectrum path: TwSet (tw_tables.cuh), NULL disables the path
};

// workspace sizes (elements) for B signals of npad matrices of degree deg0
static inline size_t tree_lev_elems(size_t B, size_t npad, size_t deg0)
{
    // coefficient form: 4 entries of degree deg0 per level-0 matrix; spectrum form of the general
    // 2x2 path (tree_low2g / tree_up): 4 entries * 2*degree values per matrix = 8*npad*deg0
    const size_t coef = 4 * npad * (deg0 + 1), spec = 8 * npad * deg0;
    return B * (coef > spec ? coef : spec) + 64;
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
static inline int launch_direct(const PairArgs &pa, int sym, fnftb_stream_t st)
{
    const long long total = (long long)pa.B * (pa.n_in / 2) * (sym ? 2 : 4);
    const int nt = 128;
    const unsigned grid = (unsigned)((total + nt - 1) / nt);
    if (sym)
        return launch_blocks<PairArgs, blk_pair_direct<DIN, true>, 128>(pa, grid, nt, 0, st,
                                                                        "tree_pair_direct");
    return launch_blocks<PairArgs, blk_pair_direct<DIN, false>, 128>(pa, grid, nt, 0, st,
                                                                     "tree_pair_direct");
}

template <int R>
static inline int launch_cols(const PairArgs &pa, int sym, fnftb_stream_t st)
{
    const long long total = (long long)pa.B * (pa.n_in / 2) * (sym ? 4 : 8) * pa.N2;
    const int nt = 128;
    const unsigned grid = (unsigned)((total + nt - 1) / nt);
    if (sym)
        return launch_blocks<PairArgs, blk_pair_cols<R, true>, 128>(pa, grid, nt, 0, st,
                                                                    "tree_pair_cols");
    return launch_blocks<PairArgs, blk_pair_cols<R, false>, 128>(pa, grid, nt, 0, st,
                                                                 "tree_pair_cols");
}

template <int R>
static inline int launch_combine(const PairArgs &pa, int sym, fnftb_stream_t st)
{
    const long long total = (long long)pa.B * (pa.n_in / 2) * (sym ? 2 : 4) * pa.N2;
    int nt = 128;
    if (nt > pa.N2)
        nt = pa.N2;  // a CTA must stay inside one (signal, pair, entry)
    const unsigned grid = (unsigned)(total / nt);
    if (sym)
        return launch_blocks<PairArgs, blk_pair_combine<R, true>, 128>(
            pa, grid, nt, sizeof(double) * nt, st, "tree_pair_combine");
    return launch_blocks<PairArgs, blk_pair_combine<R, false>, 128>(
        pa, grid, nt, sizeof(double) * nt, st, "tree_pair_combine");
}

// Runs all levels.  On entry level buffer 0 holds npad matrices of degree deg0
// per signal (and mx[0] their max, normally 1.0).  On return *cur_out tells which
// level buffer holds the single product matrix (degree deg0*npad, pending scale in
// mx[*cur_out]).  use_direct=0 forces the FFT path on every level (testing).
// tuning knobs (environment, read once): FNFT_B200_TREE_SMEM_N (row length of the
// in-shared-memory pair product), FNFT_B200_MAX_RADIX (16 or 8)
static inline int tree_knob(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

static inline const char *level_name(const char *base, int log2n, int R)
{
    // stable storage for "base_N<len>[_R<r>]" strings (names are compared by content)
    static char names[2][32][8][40];
    const int b = (base[10] == 'f' && base[13] == '_') ? 1 : 0;  // "tree_pair_fft_rows" vs "tree_pair_fft"
    int r = 0;
    while ((1 << r) < R)
        ++r;
    char *s = names[b][log2n & 31][r & 7];
    if (!s[0])
        snprintf(s, 40, "%s_N%d_R%d", base, 1 << log2n, R);
    return s;
}

static inline int launch_pair_fft(const PairArgs &pa, unsigned grid, int nt, size_t smem,
                                  fnftb_stream_t st, const char *name, int max_radix, int sym)
{
    if (sym) {
        if (max_radix < 8) {
            if (nt <= 256)
                return launch_blocks<PairArgs, blk_pair_fft_sym_r4, 256, 6>(pa, grid, nt, smem, st, name);
            return launch_blocks<PairArgs, blk_pair_fft_sym_r4, 1024, 1>(pa, grid, nt, smem, st, name);
        }
        if (max_radix < 16) {
            if (nt <= 256)
                return launch_blocks<PairArgs, blk_pair_fft_sym_r8, 256, 4>(pa, grid, nt, smem, st, name);
            return launch_blocks<PairArgs, blk_pair_fft_sym_r8, 512, 2>(pa, grid, nt, smem, st, name);
        }
        if (nt <= 256)
            return launch_blocks<PairArgs, blk_pair_fft_sym, 256, 2>(pa, grid, nt, smem, st, name);
        return launch_blocks<PairArgs, blk_pair_fft_sym, 512>(pa, grid, nt, smem, st, name);
    }
    if (max_radix < 16) {
        if (nt <= 256)
            return launch_blocks<PairArgs, blk_pair_fft_r8, 256, 4>(pa, grid, nt, smem, st, name);
        return launch_blocks<PairArgs, blk_pair_fft_r8, 512, 2>(pa, grid, nt, smem, st, name);
    }
    if (nt <= 256)
        return launch_blocks<PairArgs, blk_pair_fft, 256, 2>(pa, grid, nt, smem, st, name);
    return launch_blocks<PairArgs, blk_pair_fft, 512>(pa, grid, nt, smem, st, name);
}

static inline int tree_levels(const TreeWork &w, int B, int npad, int deg0, int normalize,
                              const TwTable &T, fnftb_stream_t st, int *cur_out,
                              int use_direct = 1, int smem_n = FNFTB_TREE_SMEM_N, int sym = 0,
                              int kappa = 0, int n_stop = 1)
{
    static const int knob_smem_n = tree_knob("FNFT_B200_TREE_SMEM_N", 0);
    static const int max_radix = tree_knob("FNFT_B200_MAX_RADIX", 16);
    static const int knob_cta_elems = tree_knob("FNFT_B200_TREE_CTA_ELEMS", 0);
    if (knob_smem_n > 0 && smem_n == FNFTB_TREE_SMEM_N)
        smem_n = knob_smem_n;
    int cur = 0;
    int n = npad, d = deg0;  // (callers may pass an intermediate level: n matrices of degree d)
    int rc = 0;
    while (n >= 2 && n > n_stop) {
        PairArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.in = w.lev[cur];
        pa.out = w.lev[1 - cur];
        pa.mx_in = w.mx[cur];
        pa.mx_out = w.mx[1 - cur];
        pa.W = w.W;
        pa.gbuf = w.gbuf;
        pa.colbuf = w.colbuf;
        pa.B = B;
        pa.n_in = n;
        pa.d_in = d;
        pa.normalize = normalize;
        pa.kappa = kappa;
        pa.T = T;
        const int npairs = n / 2;
        const int NA = sym ? 4 : 8;  // operand arrays per pair
        const bool direct =
            use_direct && (d == 1 || d == 2 || d == 3 || d == 4 || d == 6 || d == 8);
        if (direct) {
            dev_memset0(pa.mx_out, sizeof(double) * (size_t)B * npairs, st);
            switch (d) {
            case 1: rc = launch_direct<1>(pa, sym, st); break;
            case 2: rc = launch_direct<2>(pa, sym, st); break;
            case 3: rc = launch_direct<3>(pa, sym, st); break;
            case 4: rc = launch_direct<4>(pa, sym, st); break;
            case 6: rc = launch_direct<6>(pa, sym, st); break;
            default: rc = launch_direct<8>(pa, sym, st); break;
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
                int G = (knob_cta_elems > 0 ? knob_cta_elems / NA : 4 * smem_n / NA) / N;
                if (G < 1)
                    G = 1;
                if (G > npairs)
                    G = npairs;
                pa.G = G;
                pa.log2G = ilog2i((unsigned)G);
                pa.log2N2 = ilog2i((unsigned)N);
                pa.plan = make_fft_plan(N, max_radix);
                static const int knob_tpb = tree_knob("FNFT_B200_PAIR_DIV", 16);
                int nt = (NA * G * N) / knob_tpb;
                if (nt > (knob_tpb < 16 ? 1024 : 512))
                    nt = (knob_tpb < 16 ? 1024 : 512);
                if (nt < 64)
                    nt = 64;
                const unsigned grid = (unsigned)B * (unsigned)((npairs + G - 1) / G);
                rc = launch_pair_fft(pa, grid, nt, pair_smem_bytes(G, N, nt, sym), st,
                                     level_name("tree_pair_fft", pa.log2N2, 1), max_radix, sym);
            } else {
                pa.R = N / smem_n;
                pa.N2 = smem_n;
                pa.G = 1;
                pa.log2G = 0;
                pa.log2N2 = ilog2i((unsigned)smem_n);
                pa.plan = make_fft_plan(smem_n, max_radix);
                static const int knob_tpb2 = tree_knob("FNFT_B200_PAIR_DIV", 16);
                int nt = (NA * smem_n) / knob_tpb2;
                if (nt > (knob_tpb2 < 16 ? 1024 : 512))
                    nt = (knob_tpb2 < 16 ? 1024 : 512);
                if (nt < 64)
                    nt = 64;
                // column step as its own streaming kernel for R >= knob (default 4); for
                // small R the R/2-term fold inside the rows kernel's load is cheaper
                static const int knob_col_r = tree_knob("FNFT_B200_TREE_COL_R", 4);
                pa.use_col = (w.colbuf != nullptr && pa.wrap && pa.R >= knob_col_r && pa.R <= 32) ? 1 : 0;
                if (pa.use_col) {
                    switch (pa.R) {
                    case 2: rc = launch_cols<2>(pa, sym, st); break;
                    case 4: rc = launch_cols<4>(pa, sym, st); break;
                    case 8: rc = launch_cols<8>(pa, sym, st); break;
                    case 16: rc = launch_cols<16>(pa, sym, st); break;
                    default: rc = launch_cols<32>(pa, sym, st); break;
                    }
                    if (rc)
                        return rc;
                }
                const unsigned grid = (unsigned)B * (unsigned)npairs * (unsigned)pa.R;
                rc = launch_pair_fft(pa, grid, nt, pair_smem_bytes(1, smem_n, nt, sym), st,
                                     level_name("tree_pair_fft_rows", pa.log2N2, pa.R), max_radix, sym);
                if (rc)
                    return rc;
                dev_memset0(pa.mx_out, sizeof(double) * (size_t)B * npairs, st);
                switch (pa.R) {
                case 2: rc = launch_combine<2>(pa, sym, st); break;
                case 4: rc = launch_combine<4>(pa, sym, st); break;
                case 8: rc = launch_combine<8>(pa, sym, st); break;
                case 16: rc = launch_combine<16>(pa, sym, st); break;
                case 32: rc = launch_combine<32>(pa, sym, st); break;
                case 64: rc = launch_combine<64>(pa, sym, st); break;
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

// What blk_tree_final would consume: lets a caller that only needs the continuous spectrum of
// the first-row-only (NSE) mode skip the [B][4][deg+1] transfer matrix altogether -- the
// chirp-z fast path reads (a, b) from the level buffer directly (chirpz2.cuh, sym source).
struct TreeDeferred {
    int valid;          // 1: finalisation pending, fields below describe it
    int cur;            // level buffer holding the single matrix per signal
    int B, d_full, deg_out, normalize, sym, kappa;
    // the column pass of the last (row-split) level has not run yet: either the chirp-z fuses it with its first
    // stage (k_up_cols_cz) or tree_finish_cols launches it
    int cols_pending;
#ifndef FNFTB_EMUL
    UpArgs cols;
#endif
};

static inline int tree_finish_cols(TreeDeferred &f, fnftb_stream_t st)
{
    if (!f.cols_pending)
        return 0;
    f.cols_pending = 0;
#ifndef FNFTB_EMUL
    return up_cols_pending(f.cols, st);
#else
    (void)st;
    return 0;
#endif
}

static inline int tree_finalize(const TreeWork &w, int cur, int B, int d_full, int deg_out,
                                int normalize, cplx *tm, fnftb_stream_t st, int sym = 0, int kappa = 0)
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
    fa.sym = sym;
    fa.kappa = kappa;
    const long long tot = (long long)B * 4 * (deg_out + 1);
    return launch_blocks<FinalArgs, blk_tree_final>(fa, (unsigned)((tot + 255) / 256), 256, 0, st,
                                                    "tree_final");
}

// Degree the level-0 matrices are stored with (> deg0 for the chain schemes of
// leaf_chain.cuh, whose leaves are padded to a power-of-two degree); workspaces are sized by it.
static inline int tree_leaf_degree(int scheme, int deg0)
{
    ChainScheme cs;
    return chain_scheme_for(scheme, &cs) ? chain_padded_degree(deg0) : deg0;
}

// Full fast scattering for a batch: leaves -> tree -> [B][4][deg_out+1] + W[B].
static inline int tree_fscatter(const TreeWork &w, const cplx *q, const cplx *r, int B, int D,
                                int deg0, int rmode, int kappa, int scheme, double eps_t,
                                int normalize, cplx *tm, const TwTable &T, fnftb_stream_t st,
                                int use_direct = 1, int smem_n = FNFTB_TREE_SMEM_N,
                                TreeDeferred *defer = nullptr)
{
    const int npad = (int)next_pow2_sz((size_t)D);
    if (defer)
        defer->valid = 0;
    dev_memset0(w.W, sizeof(int) * (size_t)B, st);
    dev_memset0(w.status, sizeof(int) * (size_t)B, st);
    // first-row-only mode: NSE structure, and every level must be a wrap level
    // (degree a power of two) once the FFT path is reached
    static const int knob_sym = tree_knob("FNFT_B200_TREE_SYM", 1);
    const int sym = (knob_sym && rmode == FNFTB_R_NSE && (deg0 == 1 || deg0 == 2)) ? 1 : 0;
    static const int knob_low = tree_knob("FNFT_B200_TREE_LOW", 1);
    static const int knob_low_s = tree_knob("FNFT_B200_TREE_LOW_S", 0);
    static const int max_radix_low = tree_knob("FNFT_B200_MAX_RADIX", 16);
    int rc;
    int n_start = npad, d_start = deg0;
    const int dtree = tree_leaf_degree(scheme, deg0);
#ifndef FNFTB_EMUL
    // spectrum-carry path (tree_low2.cuh + tree_up.cuh): first-row-only mode; the low kernel
    // handles M*8/deg0 samples per CTA, the upper levels stay in "values at the roots of
    // unity" form until the last one
    static const int knob_low2 = tree_knob("FNFT_B200_TREE_LOW2", 7);  // 0 off, 6 or 7 = log2(M)
    static const int knob_up = tree_knob("FNFT_B200_TREE_UP", 1);      // 0: old upper levels
    static const int knob_up_smem = tree_knob("FNFT_B200_UP_SMEM_L2", 13);
    bool low2_done = false;
    if (knob_low && knob_low2 && sym && use_direct && w.tws != nullptr && r == nullptr) {
        int log2m = (knob_low2 == 6) ? 6 : 7;
        if (low2_samples(log2m, deg0) > npad)
            log2m = 6;
        const int S2 = low2_samples(log2m, deg0);
        if (S2 <= npad) {
            const int n_low = npad / S2;
            int l2n0 = ilog2i((unsigned)(2 * deg0 * S2));  // operand length of the first upper level
            bool spec = knob_up && n_low >= 2 && w.tt[0] != nullptr;
            for (int n = n_low, l = l2n0; spec && n >= 2; n /= 2, ++l)
                spec = up_supported(l, knob_up_smem);
            Low2Args lo;
            memset(&lo, 0, sizeof(lo));
            lo.q = q;
            lo.out = w.lev[0];
            lo.mx_out = w.mx[0];
            lo.W = w.W;
            lo.status = w.status;
            lo.tw = *(const TwSet *)w.tws;
            lo.tt_out = (Low2Tops *)w.tt[0];
            lo.spec_out = spec ? 1 : 0;
            lo.B = B;
            lo.D = D;
            lo.npad = npad;
            lo.kappa = kappa;
            lo.scheme = scheme;
            lo.normalize = normalize;
            lo.eps_t = eps_t;
            rc = low2_launch(lo, log2m, deg0, st);
            if (rc)
                return rc;
            n_start = n_low;
            d_start = deg0 * S2;
            low2_done = true;
            if (spec) {
                int cur = 0;
                for (int n = n_low, l = l2n0; n >= 2; n /= 2, ++l) {
                    UpArgs ua;
                    memset(&ua, 0, sizeof(ua));
                    ua.in = w.lev[cur];
                    ua.out = w.lev[1 - cur];
                    ua.tt_in = w.tt[cur];
                    ua.tt_out = w.tt[1 - cur];
                    ua.mx_in = w.mx[cur];
                    ua.mx_out = w.mx[1 - cur];
                    ua.W = w.W;
                    ua.ws = w.gbuf;
                    ua.B = B;
                    ua.n_in = n;
                    ua.l2n = l;
                    ua.normalize = normalize;
                    ua.kappa = kappa;
                    ua.last = (n == 2) ? 1 : 0;
                    ua.tw = *(const TwSet *)w.tws;
                    UpArgs pend;
                    memset(&pend, 0, sizeof(pend));
                    static const int knob_fuse = tree_knob("FNFT_B200_FUSE_COLS_CZ", 1);
                    rc = up_level(ua, knob_up_smem, st, true, (defer && ua.last && knob_fuse) ? &pend : nullptr);
                    if (rc)
                        return rc;
                    if (defer) {
                        defer->cols_pending = (pend.B != 0) ? 1 : 0;
                        defer->cols = pend;
                    }
                    cur = 1 - cur;
                }
                if (defer) {  // first-row-only result stays in the level buffer
                    defer->valid = 1;
                    defer->cur = cur;
                    defer->B = B;
                    defer->d_full = deg0 * npad;
                    defer->deg_out = deg0 * D;
                    defer->normalize = normalize;
                    defer->sym = sym;
                    defer->kappa = kappa;
                    return 0;
                }
                return tree_finalize(w, cur, B, deg0 * npad, deg0 * D, normalize, tm, st, sym, kappa);
            }
        }
    }
    // general 2x2 spectrum-carry path (tree_low2g.cuh + tree_up.cuh with E = 4): KdV, explicit r
    static const int knob_gen = tree_knob("FNFT_B200_TREE_GEN2", 1);
    if (!low2_done && !sym && knob_low && knob_gen && use_direct && w.tws != nullptr && w.tt[0] != nullptr &&
        (deg0 == 1 || deg0 == 2)) {
        const int log2m = 8;
        const int S2 = low2g_samples(log2m, deg0);
        if (S2 <= npad) {
            const int n_low = npad / S2;
            const int l2n0 = ilog2i((unsigned)(2 * deg0 * S2));
            bool spec = n_low >= 2;
            for (int n = n_low, l = l2n0; spec && n >= 2; n /= 2, ++l)
                spec = up_supported(l, knob_up_smem);
            if (spec || n_low == 1) {
                Low2gArgs lo;
                memset(&lo, 0, sizeof(lo));
                lo.q = q;
                lo.r = r;
                lo.out = w.lev[0];
                lo.mx_out = w.mx[0];
                lo.W = w.W;
                lo.status = w.status;
                lo.tt_out = (GenTops *)w.tt[0];
                lo.tw = *(const TwSet *)w.tws;
                lo.spec_out = spec ? 1 : 0;
                lo.B = B;
                lo.D = D;
                lo.npad = npad;
                lo.rmode = rmode;
                lo.kappa = kappa;
                lo.scheme = scheme;
                lo.normalize = normalize;
                lo.eps_t = eps_t;
                rc = low2g_launch(lo, deg0, st);
                if (rc)
                    return rc;
                int cur = 0;
                for (int n = n_low, l = l2n0; spec && n >= 2; n /= 2, ++l) {
                    UpArgs ua;
                    memset(&ua, 0, sizeof(ua));
                    ua.in = w.lev[cur];
                    ua.out = w.lev[1 - cur];
                    ua.tt_in = w.tt[cur];
                    ua.tt_out = w.tt[1 - cur];
                    ua.mx_in = w.mx[cur];
                    ua.mx_out = w.mx[1 - cur];
                    ua.W = w.W;
                    ua.ws = w.gbuf;
                    ua.B = B;
                    ua.n_in = n;
                    ua.l2n = l;
                    ua.normalize = normalize;
                    ua.kappa = kappa;
                    ua.last = (n == 2) ? 1 : 0;
                    ua.tw = *(const TwSet *)w.tws;
                    rc = up_level(ua, knob_up_smem, st, false);
                    if (rc)
                        return rc;
                    cur = 1 - cur;
                }
                return tree_finalize(w, cur, B, deg0 * npad, deg0 * D, normalize, tm, st, 0, kappa);
            }
        }
    }
    if (low2_done) {
    } else
#endif
    if (knob_low && (deg0 == 1 || deg0 == 2) && use_direct) {
        // fused low levels: blocks of S samples -> one matrix of degree deg0*S per block
        int S = knob_low_s > 0 ? knob_low_s : (sym ? 256 : 128);
        if (S > npad)
            S = npad;
        LowArgs lo;
        memset(&lo, 0, sizeof(lo));
        lo.q = q;
        lo.r = r;
        lo.out = w.lev[0];
        lo.mx_out = w.mx[0];
        lo.W = w.W;
        lo.status = w.status;
        lo.B = B;
        lo.D = D;
        lo.npad = npad;
        lo.deg0 = deg0;
        lo.rmode = rmode;
        lo.kappa = kappa;
        lo.scheme = scheme;
        lo.normalize = normalize;
        lo.S = S;
        lo.log2S = ilog2i((unsigned)S);
        lo.eps_t = eps_t;
        lo.T = T;
        const unsigned grid = (unsigned)B * (unsigned)(npad / S);
        const size_t smem = low_smem_bytes(sym, S, deg0);
        static const int knob_low_nt = tree_knob("FNFT_B200_LOW_NT", 128);
        const int nt = knob_low_nt;
        if (sym && max_radix_low < 8) {
            rc = launch_blocks<LowArgs, blk_tree_low_sym_r4, 512, 2>(lo, grid, nt, smem, st, "tree_low");
        } else if (sym) {
            if (max_radix_low < 16)
                rc = launch_blocks<LowArgs, blk_tree_low_sym_r8, 128, 4>(lo, grid, nt, smem, st, "tree_low");
            else
                rc = launch_blocks<LowArgs, blk_tree_low_sym, 128, 3>(lo, grid, nt, smem, st, "tree_low");
        } else {
            if (max_radix_low < 16)
                rc = launch_blocks<LowArgs, blk_tree_low_gen_r8, 128, 4>(lo, grid, nt, smem, st, "tree_low");
            else
                rc = launch_blocks<LowArgs, blk_tree_low_gen, 128, 3>(lo, grid, nt, smem, st, "tree_low");
        }
        if (rc)
            return rc;
        n_start = npad / S;
        d_start = deg0 * S;
    } else {
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
        la.sym = sym;
        la.eps_t = eps_t;
        la.status = w.status;
        const long long total = (long long)B * npad;
        LeafChainArgs ca;
        if (chain_scheme_for(scheme, &ca.cs)) {
            // higher-order splittings (degree >= 3): generated from their chains, stored with
            // a power-of-two degree (leaf_chain.cuh)
            if (ca.cs.deg != deg0)
                return -78;
            ca.la = la;
            ca.dpad = dtree;
            d_start = dtree;
            rc = launch_blocks<LeafChainArgs, blk_leaf_chain>(ca, (unsigned)((total + 127) / 128), 128,
                                                              0, st, "tree_leaf_chain");
        } else
            rc = launch_blocks<LeafArgs, blk_leaf>(la, (unsigned)((total + 127) / 128), 128, 0, st,
                                                   "tree_leaf");
        if (rc)
            return rc;
    }
    int cur = 0;
#ifndef FNFTB_EMUL
    // Long products of the general 2x2 path: coefficient levels up to degree 1024, then the
    // spectrum-carry upper levels (tree_convert.cuh).  Needs a power-of-two degree at every level.
    static const int knob_conv = tree_knob("FNFT_B200_TREE_CONVERT", 1);
    const long long total_deg = (long long)dtree * npad;
    if (knob_conv && !sym && use_direct && w.tws != nullptr && w.tt[0] != nullptr && (d_start & (d_start - 1)) == 0 &&
        d_start <= 1024 && total_deg >= 4096 && total_deg <= (1LL << 18)) {
        const int n_conv = (int)(total_deg / 1024);  // matrices of degree 1024 per signal
        bool ok = true;
        for (int n = n_conv, l = 11; ok && n >= 2; n /= 2, ++l)
            ok = up_supported(l, 13);
        if (ok) {
            rc = tree_levels(w, B, n_start, d_start, normalize, T, st, &cur, use_direct, smem_n, 0, kappa, n_conv);
            if (rc)
                return rc;
            ConvArgs ca;
            ca.in = w.lev[cur];
            ca.out = w.lev[1 - cur];
            ca.tt_out = (GenTops *)w.tt[1 - cur];
            ca.tw = *(const TwSet *)w.tws;
            ca.nmat = (long long)B * n_conv;
            rc = coef_to_spec2048(ca, st);
            if (rc)
                return rc;
            cudaMemcpyAsync(w.mx[1 - cur], w.mx[cur], sizeof(double) * (size_t)B * n_conv, cudaMemcpyDeviceToDevice, st);
            cur = 1 - cur;
            for (int n = n_conv, l = 11; n >= 2; n /= 2, ++l) {
                UpArgs ua;
                memset(&ua, 0, sizeof(ua));
                ua.in = w.lev[cur];
                ua.out = w.lev[1 - cur];
                ua.tt_in = w.tt[cur];
                ua.tt_out = w.tt[1 - cur];
                ua.mx_in = w.mx[cur];
                ua.mx_out = w.mx[1 - cur];
                ua.W = w.W;
                ua.ws = w.gbuf;
                ua.B = B;
                ua.n_in = n;
                ua.l2n = l;
                ua.normalize = normalize;
                ua.kappa = kappa;
                ua.last = (n == 2) ? 1 : 0;
                ua.tw = *(const TwSet *)w.tws;
                rc = up_level(ua, 13, st, false);
                if (rc)
                    return rc;
                cur = 1 - cur;
            }
            return tree_finalize(w, cur, B, dtree * npad, deg0 * D, normalize, tm, st, 0, kappa);
        }
    }
#endif
    rc = tree_levels(w, B, n_start, d_start, normalize, T, st, &cur, use_direct, smem_n, sym, kappa);
    if (rc)
        return rc;
    return tree_finalize(w, cur, B, dtree * npad, deg0 * D, normalize, tm, st, sym, kappa);
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

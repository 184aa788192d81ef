// fnft_b200 -- UPPER levels of the product tree in the first-row-only (NSE) mode, in the
// "spectrum carry" representation of tree_low2.cuh: a level holds its matrices (a, b) as
// VALUES at the N-th roots of unity (N = 2 * degree, bit-reversed order); a pair product
// (fnft__poly_fmult.c:239-328) is a pointwise 2x2 product giving the even bins of the next
// level's length-2N spectrum, and the odd bins are FFT_N(c_i * w_2N^i) - c_N with
// c = IFFT_N(product values).  Coefficients only exist in registers, where max|c| is
// taken for the lazy power-of-two normalisation (fnft__poly_fmult.c:330-374, see
// tree_kernels.cuh header).
//
//  k_up_smem   : N <= 8192.  CTA = (signal, pair, output polynomial): pointwise product
//                straight from HBM (+ store of the even bins), IFFT_N / twist / FFT_N in
//                shared memory, store of the odd bins.  4N cplx read (the pair's two CTAs
//                share them through L2), 2N written.
//  k_up_rows_a / k_up_cols / k_up_rows_c : longer N = R * N2 (N2 = 4096).  (a) pointwise
//                product + the inverse passes inside rows of N2 contiguous positions, (cols)
//                radix-R inverse pass across rows + coefficient fix-up + twist + radix-R
//                forward pass (streaming), (c) forward passes inside rows + odd-bin store.
//  last level  : the twist / forward half is replaced by writing the coefficients (level
//                buffer layout of tree_kernels.cuh, E = 2) for blk_tree_final.
#pragma once
#ifndef FNFTB_EMUL
#include "tree_low2.cuh"
#include "tree_low2g.cuh"

#define FNFTB_UP_ROW_L2 12  // row-split levels use rows of 4096 positions
#ifndef FNFTB_UP_PIPE
// 1: software-pipelined operand loads in the X stage (up_x_stage_pipe).  Measured on B200 (round 2,
// profiles/r02_up_experiments.md): no gain -- ptxas spreads the in-flight loads over 3 of the 6 scoreboards of
// a warp, so waiting for the oldest sub-batch also waits for the youngest one -- and k_up_rows_a gets slower.
#define FNFTB_UP_PIPE 0
#endif
#ifndef FNFTB_UP_PIPE_G
#define FNFTB_UP_PIPE_G 2      // positions per sub-batch (4 G loads)
#endif
#ifndef FNFTB_UP_PIPE_DEPTH
#define FNFTB_UP_PIPE_DEPTH 3  // sub-batches in flight
#endif

struct UpArgs {
    const cplx *in;         // [B][n_in][E][N]      (E = 2 first-row-only, 4 general)
    cplx *out;              // [B][n_in/2][E][2N], or coefficients [B][n_in/2][E][N+1] (last)
    const void *tt_in;      // [B][n_in]   Low2Tops (E = 2) or GenTops (E = 4)
    void *tt_out;           // [B][n_in/2]
    const double *mx_in;    // [B][n_in]
    double *mx_out;         // [B][n_in/2], zeroed before the level
    int *W;                 // [B]
    cplx *ws;               // [B][n_in/2][2][N] (row-split levels)
    int B, n_in, l2n;
    int normalize, kappa, last;
    int l2row;              // row-split: log2(N2)
    int pf;                 // L2 prefetch distance in CTAs (0: off), see up_prefetch
    int rows_pending;       // a pending last level (up_level with pending_cols): k_up_rows_a has not run either
    int pf_rows;            // its prefetch distance
    TwSet tw;
};

// L2 PREFETCH (round 2).  The upper levels are bound by memory-level parallelism, not by a saturated unit
// (profiles/r02_up_experiments.md): 12 warps per SM whose registers hold the loads in flight, and shared memory
// is full of work buffers, so there is no room to land asynchronous copies.  cp.async.bulk.prefetch.L2 (SASS
// UBLKPF.L2, executed by the TMA unit) needs neither registers nor shared memory: every CTA asks for the operands
// of the CTA that runs `pf` CTAs after it -- a fraction of a CTA lifetime ahead, since CTAs are dispatched in
// index order -- so that the demand loads of the X stage find their lines in L2 (~250 cycles instead of a
// loaded-DRAM latency of 800+), and the DRAM transfers overlap the compute phases of the resident CTAs.
DEV void l2_prefetch(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// n complex values starting at p, split over the first threads of the CTA in pieces of 1024 values (16 KB);
// slot0 = first thread slot to use, returns the next free slot
DEV int l2_prefetch_span(const cplx *p, int n, int tid, int slot0)
{
    const int pieces = (n + 1023) >> 10;
    const int k = tid - slot0;
    if (k >= 0 && k < pieces) {
        const int cnt = (n - (k << 10)) < 1024 ? (n - (k << 10)) : 1024;
        l2_prefetch(p + ((size_t)k << 10), (unsigned)cnt * (unsigned)sizeof(cplx));
    }
    return slot0 + pieces;
}

// SYM = first-row-only NSE mode (2 stored entries, conj symmetry), !SYM = general 2x2
template <bool SYM>
struct UpT {
    typedef Low2Tops Tops;
    static constexpr int E = 2;
};
template <>
struct UpT<false> {
    typedef GenTops Tops;
    static constexpr int E = 4;
};
DEV cplx up_top(const Low2Tops &T, int which) { return which ? T.tb : T.ta; }
DEV cplx up_bot(const Low2Tops &T, int which) { return which ? T.bb : T.ba; }
DEV cplx up_top(const GenTops &T, int which) { return T.t[which]; }
DEV cplx up_bot(const GenTops &T, int which) { return T.b[which]; }
DEV Low2Tops up_tops_product(const Low2Tops &A, const Low2Tops &B, double kap, double f)
{
    return low2_pair_tops(A, B, kap, f);
}
DEV GenTops up_tops_product(const GenTops &A, const GenTops &B, double, double f)
{
    return gen_tops_scaled(gen_pair_tops(A, B), f);
}

template <bool SYM>
struct UpPair {
    typename UpT<SYM>::Tops Tn;  // tops of the product (already scaled)
    double sc;                   // 2^-(eA+eB)
};

// per-CTA prologue: scale of the pair, tops of the product; the `leader` CTA publishes them
template <bool SYM>
DEV UpPair<SYM> up_pair_setup(const UpArgs &a, size_t sp, int s, bool leader)
{
    typedef typename UpT<SYM>::Tops Tops;
    UpPair<SYM> r;
    const size_t mA = 2 * sp, mB = 2 * sp + 1;
    int e = 0;
    if (a.normalize)
        e = rescale_exponent(a.mx_in[mA]) + rescale_exponent(a.mx_in[mB]);
    r.sc = ldexp(1.0, -e);
    const Tops TA = ((const Tops *)a.tt_in)[mA], TB = ((const Tops *)a.tt_in)[mB];
    r.Tn = up_tops_product(TA, TB, (double)a.kappa, r.sc);
    if (leader && threadIdx.x == 0) {
        ((Tops *)a.tt_out)[sp] = r.Tn;
        if (e != 0)
            atomicAdd(&a.W[s], e);
    }
    return r;
}

// generic radix-R pass at stride 1 << l2s over an array of n elements in shared memory
template <int R, int DIR>
DEV void up_p_pass(cplx *S, int n, int l2s, const TwSet &tw, int tid, int nt)
{
    constexpr int LR = Log2R<R>::value;
    const int s = 1 << l2s;
    const cplx *pt = tw.base + tw.pass_off[l2s + LR][LR];
#pragma unroll 1
    for (int idx = tid; idx < (n >> LR); idx += nt) {
        const int o = idx & (s - 1);
        const int g = idx >> l2s;
        const int base = (g << (l2s + LR)) + o;
        cplx v[R];
        if (DIR > 0) {
#pragma unroll
            for (int q = 0; q < R; ++q)
                v[q] = S[swz2(base + (brev_c(q, LR) << l2s))];
            up_twiddle_mul<R, true>(v, pt, s, o);
            Dft<R, +1>::run(v);
#pragma unroll
            for (int n2 = 0; n2 < R; ++n2)
                S[swz2(base + (n2 << l2s))] = v[n2];
        } else {
#pragma unroll
            for (int n2 = 0; n2 < R; ++n2)
                v[n2] = S[swz2(base + (n2 << l2s))];
            Dft<R, -1>::run(v);
            up_twiddle_mul<R, false>(v, pt, s, o);
#pragma unroll
            for (int q = 0; q < R; ++q)
                S[swz2(base + (brev_c(q, LR) << l2s))] = v[q];
        }
    }
}

// X stage: pointwise product of the CTA's positions [p0, p0 + n) (global), store of the even
// bins, first inverse pass (radix RX, stride 1) into shared memory (local index l0..l0+n).
// A warp owns chunks of 32*RX consecutive positions: lanes run over consecutive positions for
// the global loads / stores (512 contiguous bytes per access) and park the product values in
// shared memory; after a __syncwarp each lane takes its RX consecutive positions for the
// stride-1 pass.
template <int RX, bool SYM>
DEV void up_x_stage(const UpArgs &a, size_t sp, int which, int p0, int l0, int n, double sg,
                    const UpPair<SYM> &P, cplx *S, int tid, int nt)
{
    constexpr int LR = Log2R<RX>::value;
    constexpr int E = UpT<SYM>::E;
    const int N = 1 << a.l2n;
    const cplx *mA = a.in + (2 * E * sp) * (size_t)N;  // matrix A, then matrix B
    const cplx *mB = mA + (size_t)E * N;
    // SYM:  which = 0: c = aA*aB + (-kap*sg*bA)*conj(bB);  which = 1: c = aA*bB + (sg*bA)*conj(aB)
    // !SYM: c(row, col) = A(row,1)*B(1,col) + A(row,2)*B(2,col)
    const cplx *aA = SYM ? mA : mA + (size_t)(2 * (which >> 1)) * N;
    const cplx *bA = aA + N;
    const cplx *pz = SYM ? mB + (which ? 1 : 0) * (size_t)N : mB + (size_t)(which & 1) * N;
    const cplx *pw = SYM ? mB + (which ? 0 : 1) * (size_t)N : mB + (size_t)(2 + (which & 1)) * N;
    cplx *ge = a.last ? nullptr : a.out + (E * sp + which) * (size_t)(2 * N);
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const double sc = P.sc;
    const double f = SYM ? (which ? sg * sc : -(double)a.kappa * sg * sc) : sc;
#pragma unroll 1
    for (int cb = warp * (32 * RX); cb < n; cb += nwarps * (32 * RX)) {
#pragma unroll
        for (int i0 = 0; i0 < RX; i0 += 4) {
            cplx x[4], y[4], z[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pos = p0 + cb + 32 * (i0 + i) + lane;
                x[i] = __ldg(&aA[pos]);
                y[i] = __ldg(&bA[pos]);
                z[i] = __ldg(&pz[pos]);
                w[i] = __ldg(&pw[pos]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int loc = cb + 32 * (i0 + i) + lane;
                cplx r = cmul(cscale(x[i], sc), z[i]);
                if (SYM)
                    cfmac(r, cscale(y[i], f), w[i]);
                else
                    cfma(r, cscale(y[i], f), w[i]);
                if (ge)
                    ge[p0 + loc] = r;
                S[swz2(l0 + loc)] = r;
            }
        }
        __syncwarp();
        const int b = l0 + cb + RX * lane;
        cplx v[RX];
#pragma unroll
        for (int q = 0; q < RX; ++q)
            v[q] = S[swz2(b + brev_c(q, LR))];
        Dft<RX, +1>::run(v);
#pragma unroll
        for (int j = 0; j < RX; ++j)
            S[swz2(b + j)] = v[j];
    }
}

// prefetch of the four operand streams that up_x_stage reads for (pair sp, entry which), positions [p0, p0 + n)
template <bool SYM>
DEV void up_prefetch_operands(const UpArgs &a, size_t sp, int which, int p0, int n, int tid)
{
    constexpr int E = UpT<SYM>::E;
    const int N = 1 << a.l2n;
    const cplx *mA = a.in + (2 * E * sp) * (size_t)N;
    const cplx *mB = mA + (size_t)E * N;
    if (SYM) {
        if (n == N) {  // whole pair: one contiguous block [aA bA aB bB]
            l2_prefetch_span(mA, 4 * N, tid, 0);
            return;
        }
        int sl = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            sl = l2_prefetch_span(mA + (size_t)e * N + p0, n, tid, sl);
    } else {
        const cplx *aA = mA + (size_t)(2 * (which >> 1)) * N;
        int sl = l2_prefetch_span(aA + p0, n, tid, 0);
        sl = l2_prefetch_span(aA + N + p0, n, tid, sl);
        sl = l2_prefetch_span(mB + (size_t)(which & 1) * N + p0, n, tid, sl);
        l2_prefetch_span(mB + (size_t)(2 + (which & 1)) * N + p0, n, tid, sl);
    }
}

// streaming operand load: read-only path, no L1 allocation (the twiddle tables stay in L1)
DEV cplx up_ld_stream(const cplx *p)
{
    cplx v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// X stage with SOFTWARE-PIPELINED operand loads (round 2).  ncu, round 1: 66 % of the warp samples of
// k_up_smem<12> sat in the X stage on long_scoreboard -- every lane loaded 16 values, waited for all of them,
// computed, and only then issued the next 16, so on average half of the possible bytes were in flight.  Here
// the NPOS positions of a lane are split into sub-batches of G positions (4 G loads); DEPTH sub-batches are
// in flight while the oldest one is consumed.  The position range [p0, p0 + n) may straddle the half
// boundary hb of the spectrum ((-1)^k changes sign there); chunks of 32*RX positions never do.
// The pair prologue (dependent loads of max|c| and the tops) runs after the first loads have been issued.
template <int RX, bool SYM, int NPOS, int G, int DEPTH>
DEV UpPair<SYM> up_x_stage_pipe(const UpArgs &a, size_t sp, int s, int which, bool leader, int p0, int l0, int hb,
                                double sg0, cplx *S, int tid, int nt)
{
    constexpr int LR = Log2R<RX>::value;
    constexpr int E = UpT<SYM>::E;
    constexpr int NSB = NPOS / G;       // sub-batches per lane
    constexpr int SBC = RX / G;         // sub-batches per chunk of 32*RX positions
    static_assert(NPOS % RX == 0 && RX % G == 0 && DEPTH >= 2 && DEPTH <= NSB, "shape");
    const int N = 1 << a.l2n;
    const cplx *mA = a.in + (2 * E * sp) * (size_t)N;  // matrix A, then matrix B
    const cplx *mB = mA + (size_t)E * N;
    const cplx *aA = SYM ? mA : mA + (size_t)(2 * (which >> 1)) * N;
    const cplx *bA = aA + N;
    const cplx *pz = SYM ? mB + (which ? 1 : 0) * (size_t)N : mB + (size_t)(which & 1) * N;
    const cplx *pw = SYM ? mB + (which ? 0 : 1) * (size_t)N : mB + (size_t)(2 + (which & 1)) * N;
    cplx *ge = a.last ? nullptr : a.out + (E * sp + which) * (size_t)(2 * N);
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    // local offset of position i of sub-batch sb
    auto off = [&](int sb, int i) {
        return ((sb / SBC) * nwarps + warp) * (32 * RX) + 32 * ((sb % SBC) * G + i) + lane;
    };
    cplx bx[DEPTH][G], by[DEPTH][G], bz[DEPTH][G], bw[DEPTH][G];
    auto issue = [&](int sb, int slot) {
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int pos = p0 + off(sb, i);
            bx[slot][i] = up_ld_stream(&aA[pos]);
            by[slot][i] = up_ld_stream(&bA[pos]);
            bz[slot][i] = up_ld_stream(&pz[pos]);
            bw[slot][i] = up_ld_stream(&pw[pos]);
        }
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d)
        issue(d, d);
    const UpPair<SYM> P = up_pair_setup<SYM>(a, sp, s, leader);
    const double sc = P.sc;
#pragma unroll
    for (int sb = 0; sb < NSB; ++sb) {
        if (sb + DEPTH - 1 < NSB)
            issue(sb + DEPTH - 1, (sb + DEPTH - 1) % DEPTH);
        const int slot = sb % DEPTH;
        const double sg = (p0 + off(sb, 0) >= hb) ? -sg0 : sg0;
        const double f = SYM ? (which ? sg * sc : -(double)a.kappa * sg * sc) : sc;
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int loc = off(sb, i);
            cplx r = cmul(cscale(bx[slot][i], sc), bz[slot][i]);
            if (SYM)
                cfmac(r, cscale(by[slot][i], f), bw[slot][i]);
            else
                cfma(r, cscale(by[slot][i], f), bw[slot][i]);
            if (ge)
                ge[p0 + loc] = r;
            S[swz2(l0 + loc)] = r;
        }
        if (sb % SBC == SBC - 1) {  // chunk complete: first inverse pass (radix RX, stride 1) on it
            __syncwarp();
            const int b = l0 + ((sb / SBC) * nwarps + warp) * (32 * RX) + RX * lane;
            cplx v[RX];
#pragma unroll
            for (int q = 0; q < RX; ++q)
                v[q] = S[swz2(b + brev_c(q, LR))];
            Dft<RX, +1>::run(v);
#pragma unroll
            for (int j = 0; j < RX; ++j)
                S[swz2(b + j)] = v[j];
        }
    }
    return P;
}

// F stage: last forward pass (radix RX, stride 1) of the CTA's n local elements, "- c_N",
// coalesced store to the odd-bin region starting at godd (same warp-chunk scheme)
template <int RX>
DEV void up_f_stage(cplx *S, int n, cplx ct, cplx *godd, int tid, int nt)
{
    constexpr int LR = Log2R<RX>::value;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
#pragma unroll 1
    for (int cb = warp * (32 * RX); cb < n; cb += nwarps * (32 * RX)) {
        const int b = cb + RX * lane;
        cplx v[RX];
#pragma unroll
        for (int j = 0; j < RX; ++j)
            v[j] = S[swz2(b + j)];
        Dft<RX, -1>::run(v);
#pragma unroll
        for (int q = 0; q < RX; ++q)
            S[swz2(b + brev_c(q, LR))] = v[q];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RX; ++i) {
            const int loc = cb + 32 * i + lane;
            godd[loc] = csub(S[swz2(loc)], ct);
        }
    }
}

// M stage (radix 16, stride N/16) on a whole length-N array in shared memory; `last` writes
// the coefficients to gcoef instead of twisting and transforming forward.  Returns max|c|^2.
template <bool SYM>
DEV double up_m_stage(cplx *S, int l2n, const TwSet &tw, const UpPair<SYM> &P, int which, bool last,
                      cplx *gcoef, int tid, int nt)
{
    constexpr int R = 16, LR = 4;
    const int l2s = l2n - LR;
    const int s = 1 << l2s;
    const int N = 1 << l2n;
    const double invN = 1.0 / (double)N;
    const cplx *pt = tw.base + tw.pass_off[l2n][LR];
    const cplx *tt = tw.base + tw.twist_off[l2n];
    double m2 = 0.0;
#pragma unroll 1
    for (int o = tid; o < s; o += nt) {
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q)
            v[q] = S[swz2(o + (brev_c(q, LR) << l2s))];
        up_twiddle_mul<R, true>(v, pt, s, o);
        Dft<R, +1>::run(v);
#pragma unroll
        for (int n = 0; n < R; ++n)
            v[n] = cscale(v[n], invN);
        if (o == 0)
            v[0] = up_bot(P.Tn, which);
#pragma unroll
        for (int n = 0; n < R; ++n)
            m2 = fmax(m2, cabs2(v[n]));
        if (last) {
#pragma unroll
            for (int n = 0; n < R; ++n)
                gcoef[o + (n << l2s)] = v[n];
            if (o == 0)
                gcoef[N] = up_top(P.Tn, which);
            continue;
        }
#if FNFTB_TW_DERIVE
        {
            // twist w_2N^(o + n*s) = w_2N^o * w_2R^n: one load, the second factor is a constant
            const cplx t0 = __ldg(&tt[o]);
            up_twist16(v, t0);
        }
#else
#pragma unroll
        for (int n0 = 0; n0 < R; n0 += 8) {
            cplx w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                w[j] = __ldg(&tt[o + ((n0 + j) << l2s)]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[n0 + j] = cmul(v[n0 + j], w[j]);
        }
#endif
        Dft<R, -1>::run(v);
        up_twiddle_mul<R, false>(v, pt, s, o);
#pragma unroll
        for (int q = 0; q < R; ++q)
            S[swz2(o + (brev_c(q, LR) << l2s))] = v[q];
    }
    return m2;
}

// publishes max(|c|) of a CTA: warp shuffle, shared memory, one atomic
DEV void up_publish_max(double m2, double *red, double *dst, int tid, int nt)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, off));
    if ((tid & 31) == 0)
        red[tid >> 5] = m2;
    __syncthreads();
    if (tid == 0) {
        double m = red[0];
        for (int w = 1; w < nt / 32; ++w)
            m = fmax(m, red[w]);
        atomic_max_double(dst, sqrt(m));
    }
}

// inverse / forward row passes between the stride-1 stage (radix RX) and length 1 << l2len
//   l2len 11: RX = 8,  P16(s=8)            [+ M16 at s=128]
//   l2len 12: RX = 16, P16(s=16)           [+ M16 at s=256]
//   l2len 13: RX = 8,  P8(s=8), P8(s=64)   [+ M16 at s=512]
// with_top: rows of a longer transform also run the pass that the M stage replaces
template <int DIR, int L2LEN, bool WITH_TOP>
DEV void up_row_passes(cplx *S, const TwSet &tw, int tid, int nt)
{
    constexpr int n = 1 << L2LEN;
    if constexpr (DIR > 0) {
        if constexpr (L2LEN == 12) {
            up_p_pass<16, +1>(S, n, 4, tw, tid, nt);
            __syncthreads();
        } else if constexpr (L2LEN == 11) {
            up_p_pass<16, +1>(S, n, 3, tw, tid, nt);
            __syncthreads();
        } else {
            up_p_pass<8, +1>(S, n, 3, tw, tid, nt);
            __syncthreads();
            up_p_pass<8, +1>(S, n, 6, tw, tid, nt);
            __syncthreads();
        }
        if constexpr (WITH_TOP) {
            up_p_pass<16, +1>(S, n, L2LEN - 4, tw, tid, nt);
            __syncthreads();
        }
    } else {
        if constexpr (WITH_TOP) {
            up_p_pass<16, -1>(S, n, L2LEN - 4, tw, tid, nt);
            __syncthreads();
        }
        if constexpr (L2LEN == 12) {
            up_p_pass<16, -1>(S, n, 4, tw, tid, nt);
            __syncthreads();
        } else if constexpr (L2LEN == 11) {
            up_p_pass<16, -1>(S, n, 3, tw, tid, nt);
            __syncthreads();
        } else {
            up_p_pass<8, -1>(S, n, 6, tw, tid, nt);
            __syncthreads();
            up_p_pass<8, -1>(S, n, 3, tw, tid, nt);
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// whole level in shared memory: grid.x = B * npairs * 2, blockDim.x = N / 32
// ---------------------------------------------------------------------------------------
// spectrum positions per thread of k_up_smem: 32 (168 registers) for N = 2048 / 4096, 16 (128 registers, 512 threads)
// for N = 8192, whose single CTA per SM then has 16 warps (measured with the L2 prefetch on: 5.03 -> 4.84 ms per 4096
// signals; for the smaller N 16 positions per thread lose: 3.53 -> 3.73, 4.01 -> 5.20)
#ifndef FNFTB_UP_TPP
#define FNFTB_UP_TPP(L2N) (((L2N) == 13) ? 16 : 32)
#endif
template <int L2N, bool SYM>
__global__ void __launch_bounds__((1 << L2N) / FNFTB_UP_TPP(L2N), (L2N == 11) ? 6 : ((L2N == 12) ? 3 : 1))
    k_up_smem(const UpArgs a)
{
    constexpr int E = UpT<SYM>::E;
    constexpr int N = 1 << L2N, NT = N / FNFTB_UP_TPP(L2N);
    constexpr int RX = (L2N == 12) ? 16 : 8;
    extern __shared__ double2 fnftb_smem_up[];
    cplx *S = (cplx *)fnftb_smem_up;
    double *red = (double *)(S + N);
    const int tid = threadIdx.x;
    const int which = blockIdx.x % E;
    const size_t sp = blockIdx.x / E;
    const int npairs = a.n_in >> 1;
    const int s = (int)(sp / npairs);
    if (a.pf > 0) {  // operands of the CTA `pf` places ahead (the SYM pair's two CTAs read the same block)
        const size_t bp = (size_t)blockIdx.x + (size_t)a.pf;
        if (bp < gridDim.x && (!SYM || (bp % E) == 0))
            up_prefetch_operands<SYM>(a, bp / E, (int)(bp % E), 0, N, tid);
    }
#if FNFTB_UP_PIPE
    // X stage over all N positions; the half regions carry (-1)^k = +1 / -1
    const UpPair<SYM> P = up_x_stage_pipe<RX, SYM, FNFTB_UP_TPP(L2N), FNFTB_UP_PIPE_G, FNFTB_UP_PIPE_DEPTH>(a, sp, s, which, which == 0, 0, 0, N / 2,
                                                                                        1.0, S, tid, NT);
#else
    const UpPair<SYM> P = up_pair_setup<SYM>(a, sp, s, which == 0);

    // X stage: the two half regions carry (-1)^k = +1 / -1
    up_x_stage<RX, SYM>(a, sp, which, 0, 0, N / 2, 1.0, P, S, tid, NT);
    up_x_stage<RX, SYM>(a, sp, which, N / 2, N / 2, N / 2, -1.0, P, S, tid, NT);
#endif
    __syncthreads();
    up_row_passes<+1, L2N, false>(S, a.tw, tid, NT);
    cplx *gcoef = a.last ? a.out + (E * sp + which) * (size_t)(N + 1) : nullptr;
    const double m2 = up_m_stage<SYM>(S, L2N, a.tw, P, which, a.last != 0, gcoef, tid, NT);
    __syncthreads();
    if (!a.last) {
        up_row_passes<-1, L2N, false>(S, a.tw, tid, NT);
        cplx *godd = a.out + (E * sp + which) * (size_t)(2 * N) + N;
        up_f_stage<RX>(S, N, up_top(P.Tn, which), godd, tid, NT);
    }
    double mm = m2;
    if (tid == 0)
        mm = fmax(mm, cabs2(up_top(P.Tn, which)));
    up_publish_max(mm, red, &a.mx_out[sp], tid, NT);
}

// ---------------------------------------------------------------------------------------
// N = 8192 as a CLUSTER of two CTAs (round 2).  k_up_smem<13> needs the whole 128 KiB work buffer in one CTA, so one
// CTA is resident per SM and its memory phase (X stage) and compute phases never overlap anything (3.5 TB/s where
// the N = 4096 kernel with three CTAs per SM reaches 4.35).  Here each CTA of a pair holds one half of the positions
// (64 KiB, three CTAs per SM like N = 4096): the X stage, the inverse / forward passes at strides < 4096 and the
// F stage stay inside a half -- the two halves are exactly the regions with (-1)^k = +1 / -1 -- and only the M stage
// (radix 16 at stride 512: 8 elements from either half per butterfly) reads and writes the partner's half through
// distributed shared memory, with a cluster barrier on either side.
// grid.x = B * npairs * E * 2, cluster (2, 1, 1), blockDim.x = 128
// ---------------------------------------------------------------------------------------
DEV unsigned up_cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
DEV void up_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of the same shared-memory location in CTA `rank` of the cluster
DEV cplx *up_cluster_map(cplx *p, unsigned rank)
{
    unsigned long long in = (unsigned long long)p, out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(in), "r"(rank));
    return (cplx *)out;
}

// L2N = 13: cluster of 2 (RX = 8, passes at strides 8 and 64, M stage at stride 512); L2N = 14: cluster of 4 (RX = 16,
// passes at strides 16 and 128, M stage at stride 1024: four elements of a butterfly in each quarter) -- this one
// replaces the three row-split kernels of the N = 16384 level and their workspace round trip (34 GB -> 17 GB).
// L2H: log2 of the positions per CTA (12: 128 threads, three CTAs per SM; 11: 64 threads, six CTAs per SM)
template <int L2N, int L2H, bool SYM>
__global__ void __launch_bounds__((1 << L2H) / 32, (L2H == 12) ? 3 : 6) k_up_smem_cluster(const UpArgs a)
{
    constexpr int E = UpT<SYM>::E;
    constexpr int N = 1 << L2N, H = 1 << L2H, CL = N / H, NT = H / 32;
    constexpr int RX = (L2N == 13) ? 8 : 16, L2RX = (L2N == 13) ? 3 : 4;
    extern __shared__ double2 fnftb_smem_up[];
    cplx *S = (cplx *)fnftb_smem_up;
    double *red = (double *)(S + H);
    const int tid = threadIdx.x;
    const unsigned h = up_cluster_rank();
    const unsigned item = blockIdx.x / CL;
    const int which = item % E;
    const size_t sp = item / E;
    const int npairs = a.n_in >> 1;
    const int s = (int)(sp / npairs);
    if (a.pf > 0 && h == 0) {  // operands of the item `pf` places ahead, issued by the first CTA of the cluster only
        const size_t ip = (size_t)item + (size_t)a.pf;
        if (ip < (gridDim.x / CL) && (!SYM || (ip % E) == 0))
            up_prefetch_operands<SYM>(a, ip / E, (int)(ip % E), 0, N, tid);
    }
    const UpPair<SYM> P = up_pair_setup<SYM>(a, sp, s, which == 0 && h == 0);
    up_x_stage<RX, SYM>(a, sp, which, (int)h * H, 0, H, (h >= CL / 2) ? -1.0 : 1.0, P, S, tid, NT);
    __syncthreads();
    if constexpr (L2N == 12) {  // 16 * 16 * 16
        up_p_pass<16, +1>(S, H, 4, a.tw, tid, NT);
    } else {                    // 8 * 8 * 8 * 16 or 16 * 8 * 8 * 16
        up_p_pass<8, +1>(S, H, L2RX, a.tw, tid, NT);
        __syncthreads();
        up_p_pass<8, +1>(S, H, L2RX + 3, a.tw, tid, NT);
    }
    up_cluster_sync();
    // M stage: this CTA's share of the N / 16 butterflies; element o + r * st, r < 16, lives in CTA r / (16 / CL)
    cplx *Sr[CL];
#pragma unroll
    for (int c = 0; c < CL; ++c)
        Sr[c] = up_cluster_map(S, (unsigned)c);
    double m2 = 0.0;
    {
        constexpr int R = 16, LR = 4, l2s = L2N - LR, st = 1 << l2s, PER = R / CL;
        const double invN = 1.0 / (double)N;
        const cplx *pt = a.tw.base + a.tw.pass_off[L2N][LR];
        const cplx *tt = a.tw.base + a.tw.twist_off[L2N];
#pragma unroll 1
        for (int o = (int)h * (st / CL) + tid; o < ((int)h + 1) * (st / CL); o += NT) {
            cplx v[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int r = brev_c(q, LR);
                v[q] = Sr[r / PER][swz2(o + ((r % PER) << l2s))];
            }
            up_twiddle_mul<R, true>(v, pt, st, o);
            Dft<R, +1>::run(v);
#pragma unroll
            for (int n = 0; n < R; ++n)
                v[n] = cscale(v[n], invN);
            if (o == 0)
                v[0] = up_bot(P.Tn, which);
#pragma unroll
            for (int n = 0; n < R; ++n)
                m2 = fmax(m2, cabs2(v[n]));
            up_twist16(v, __ldg(&tt[o]));
            Dft<R, -1>::run(v);
            up_twiddle_mul<R, false>(v, pt, st, o);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int r = brev_c(q, LR);
                Sr[r / PER][swz2(o + ((r % PER) << l2s))] = v[q];
            }
        }
    }
    up_cluster_sync();
    if constexpr (L2N == 12) {
        up_p_pass<16, -1>(S, H, 4, a.tw, tid, NT);
    } else {
        up_p_pass<8, -1>(S, H, L2RX + 3, a.tw, tid, NT);
        __syncthreads();
        up_p_pass<8, -1>(S, H, L2RX, a.tw, tid, NT);
    }
    __syncthreads();
    cplx *godd = a.out + (E * sp + which) * (size_t)(2 * N) + N + (size_t)h * H;
    up_f_stage<RX>(S, H, up_top(P.Tn, which), godd, tid, NT);
    double mm = m2;
    if (tid == 0 && h == 0)
        mm = fmax(mm, cabs2(up_top(P.Tn, which)));
    up_publish_max(mm, red, &a.mx_out[sp], tid, NT);
}

// ---------------------------------------------------------------------------------------
// row-split levels, N = R * N2
// (a) grid.x = B * npairs * 2 * R: pointwise product + inverse passes inside row r
// ---------------------------------------------------------------------------------------
template <int NT, bool SYM>
__global__ void __launch_bounds__(NT, 3) k_up_rows_a(const UpArgs a)
{
    constexpr int E = UpT<SYM>::E;
    extern __shared__ double2 fnftb_smem_up[];
    cplx *S = (cplx *)fnftb_smem_up;
    const int l2n = a.l2n, l2row = a.l2row, N2 = 1 << l2row;
    const int l2R = l2n - l2row;
    const int tid = threadIdx.x;
    const int row = blockIdx.x & ((1 << l2R) - 1);
    const size_t arr = blockIdx.x >> l2R;
    const int which = (int)(arr % E);
    const size_t sp = arr / E;
    const int npairs = a.n_in >> 1;
    const int s = (int)(sp / npairs);
    const double sg = (row >> (l2R - 1)) ? -1.0 : 1.0;
    if (a.pf > 0) {
        const size_t bp = (size_t)blockIdx.x + (size_t)a.pf;
        const size_t arrp = bp >> l2R;
        if (bp < gridDim.x && (!SYM || (arrp % E) == 0))
            up_prefetch_operands<SYM>(a, arrp / E, (int)(arrp % E), (int)(bp & ((1 << l2R) - 1)) << l2row, N2, tid);
    }
    // X stage over positions [row*N2, (row+1)*N2): local index = position - row*N2
#if FNFTB_UP_PIPE
    static_assert(NT == 128, "32 positions per lane");
    up_x_stage_pipe<16, SYM, 32, FNFTB_UP_PIPE_G, FNFTB_UP_PIPE_DEPTH>(a, sp, s, which, which == 0 && row == 0, row << l2row, 0,
                                                                      0x7fffffff, sg, S, tid, NT);
#else
    const UpPair<SYM> P = up_pair_setup<SYM>(a, sp, s, which == 0 && row == 0);
    up_x_stage<16, SYM>(a, sp, which, row << l2row, 0, N2, sg, P, S, tid, NT);
#endif
    __syncthreads();
    up_row_passes<+1, FNFTB_UP_ROW_L2, true>(S, a.tw, tid, NT);
    cplx *dst = a.ws + arr * ((size_t)1 << l2n) + ((size_t)row << l2row);
    for (int i = tid; i < N2; i += NT)
        dst[i] = S[swz2(i)];
}

// (cols) one thread per (array, o < N2): radix-R inverse pass across the rows, coefficient
// fix-up and max, then twist + radix-R forward pass (or the coefficient output when last)
template <int R, bool SYM>
__global__ void __launch_bounds__((R >= 64) ? 128 : 256, (R >= 32) ? 1 : ((R >= 8) ? 3 : 4)) k_up_cols(const UpArgs a)
{
    typedef typename UpT<SYM>::Tops Tops;
    constexpr int E = UpT<SYM>::E;
    constexpr int LR = Log2R<R>::value;
    __shared__ double red[8];
    const int l2n = a.l2n, l2row = l2n - LR;
    const int N = 1 << l2n;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int o = (int)(gid & ((1 << l2row) - 1));
    const size_t arr = (size_t)(gid >> l2row);  // a CTA stays inside one array (256 | N2)
    const int which = (int)(arr % E);
    const size_t sp = arr / E;
    const Tops Tn = ((const Tops *)a.tt_out)[sp];
    cplx *w = a.ws + arr * (size_t)N;
    const cplx *pt = a.tw.base + a.tw.pass_off[l2n][LR];
    const cplx *tt = a.tw.base + a.tw.twist_off[l2n];
    const double invN = 1.0 / (double)N;
    if (a.pf > 0 && threadIdx.x < R) {  // the R column segments of the CTA `pf` places ahead
        const size_t bp = (size_t)blockIdx.x + (size_t)a.pf;
        if (bp < gridDim.x) {
            const long long g0 = (long long)bp * blockDim.x;
            const cplx *wp = a.ws + (size_t)(g0 >> l2row) * (size_t)N + (size_t)(g0 & ((1 << l2row) - 1));
            l2_prefetch(wp + ((size_t)brev_c((int)threadIdx.x, LR) << l2row), blockDim.x * (unsigned)sizeof(cplx));
        }
    }
    cplx v[R];
#pragma unroll
    for (int q = 0; q < R; ++q)
        v[q] = w[o + ((size_t)brev_c(q, LR) << l2row)];
    up_twiddle_mul<R, true>(v, pt, 1 << l2row, o);
    Dft<R, +1>::run(v);
#pragma unroll
    for (int n = 0; n < R; ++n)
        v[n] = cscale(v[n], invN);
    if (o == 0)
        v[0] = up_bot(Tn, which);
    double m2 = 0.0;
#pragma unroll
    for (int n = 0; n < R; ++n)
        m2 = fmax(m2, cabs2(v[n]));
    if (a.last) {
        cplx *gcoef = a.out + arr * (size_t)(N + 1);
#pragma unroll
        for (int n = 0; n < R; ++n)
            gcoef[o + ((size_t)n << l2row)] = v[n];
        if (o == 0) {
            const cplx ct = up_top(Tn, which);
            gcoef[N] = ct;
            m2 = fmax(m2, cabs2(ct));
        }
    } else {
        if (o == 0)
            m2 = fmax(m2, cabs2(up_top(Tn, which)));
        UpTwist<R, 0>::run(v, __ldg(&tt[o]));
        Dft<R, -1>::run(v);
        up_twiddle_mul<R, false>(v, pt, 1 << l2row, o);
#pragma unroll
        for (int q = 0; q < R; ++q)
            w[o + ((size_t)brev_c(q, LR) << l2row)] = v[q];
    }
    up_publish_max(m2, red, &a.mx_out[sp], threadIdx.x, blockDim.x);
}

// (c) grid.x = B * npairs * 2 * R: forward passes inside row `row` of the workspace, odd bins
template <int NT, bool SYM>
__global__ void __launch_bounds__(NT, 3) k_up_rows_c(const UpArgs a)
{
    typedef typename UpT<SYM>::Tops Tops;
    constexpr int E = UpT<SYM>::E;
    extern __shared__ double2 fnftb_smem_up[];
    cplx *S = (cplx *)fnftb_smem_up;
    const int l2n = a.l2n, l2row = a.l2row, N2 = 1 << l2row;
    const int l2R = l2n - l2row;
    const int tid = threadIdx.x;
    const int row = blockIdx.x & ((1 << l2R) - 1);
    const size_t arr = blockIdx.x >> l2R;
    const int which = (int)(arr % E);
    const size_t sp = arr / E;
    const Tops Tn = ((const Tops *)a.tt_out)[sp];
    const cplx *src = a.ws + arr * ((size_t)1 << l2n) + ((size_t)row << l2row);
    if (a.pf > 0) {
        const size_t bp = (size_t)blockIdx.x + (size_t)a.pf;
        if (bp < gridDim.x)
            l2_prefetch_span(a.ws + (bp << l2row), N2, tid, 0);  // rows are contiguous in block order
    }
    for (int i = tid; i < N2; i += NT)
        S[swz2(i)] = src[i];
    __syncthreads();
    up_row_passes<-1, FNFTB_UP_ROW_L2, true>(S, a.tw, tid, NT);
    cplx *godd = a.out + arr * ((size_t)2 << l2n) + ((size_t)1 << l2n) + ((size_t)row << l2row);
    up_f_stage<16>(S, N2, up_top(Tn, which), godd, tid, NT);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

// can this level (operand length 1 << l2n) run on the spectrum path?
static inline bool up_supported(int l2n, int l2smem_max)
{
    if (l2n < 11 || l2n > FNFTB_TW_MAXL)
        return false;
    if (l2n <= l2smem_max)
        return true;
    const int l2R = l2n - FNFTB_UP_ROW_L2;
    return l2R >= 1 && l2R <= 6;  // radix 32 / 64 across the rows for operand lengths 2^17 / 2^18
}

// kernels instantiated in k_tree_up.cu only
#ifdef FNFTB_TU_UP
template <class K>
static inline int up_launch(K kernel, const UpArgs &a, unsigned grid, int nt, size_t smem, cudaStream_t st,
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

// L2 prefetch distance in CTAs per kernel family (environment FNFT_B200_PFD_<family>; 0 = off).  Measured on
// B200 (profiles/r02_l2_prefetch.md): the best distance is about one loaded-DRAM latency of CTA starts -- the L2
// does not retain a line much longer than 10 us at 4 TB/s of streaming traffic, so longer distances lose.
enum { UP_PF_SMEM11 = 0, UP_PF_SMEM12, UP_PF_SMEM13, UP_PF_ROWS_A, UP_PF_COLS, UP_PF_ROWS_C, UP_PF_FAMILIES };
static inline int up_pf_distance(int family)
{
    static int init = 0, dist[UP_PF_FAMILIES];
    static const char *names[UP_PF_FAMILIES] = {"FNFT_B200_PFD_SMEM11", "FNFT_B200_PFD_SMEM12", "FNFT_B200_PFD_SMEM13",
                                                "FNFT_B200_PFD_ROWS_A", "FNFT_B200_PFD_COLS",   "FNFT_B200_PFD_ROWS_C"};
    static const int dflt[UP_PF_FAMILIES] = {8, 4, 2, 2, 6, 6};
    if (!init) {
        for (int f = 0; f < UP_PF_FAMILIES; ++f) {
            const char *e = getenv(names[f]);
            dist[f] = (e && e[0]) ? atoi(e) : dflt[f];
        }
        init = 1;
    }
    return dist[family];
}

// cluster launch of k_up_smem_cluster: `items` work items, CL CTAs of 128 threads and 64 KiB each per item
template <class K>
static inline int up_launch_cluster(K kern, const UpArgs &a, unsigned items, unsigned CL, cudaStream_t st,
                                    const char *name, unsigned H = 4096)
{
    const size_t smem_h = sizeof(cplx) * H + 64 * sizeof(double);
    const int eo = fnftb_smem_optin((const void *)kern, smem_h);
    if (eo != 0)
        return eo;
    cudaError_t e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(items * CL, 1, 1);
    cfg.blockDim = dim3(H / 32, 1, 1);
    cfg.dynamicSmemBytes = smem_h;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (g_fnftb_profile_on)
        fnftb_profile_begin(name, st);
    e = cudaLaunchKernelEx(&cfg, kern, a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(st);
    ++g_fnftb_launch_count;
    return (int)e;
}

// launches the column pass of a row-split level (a.l2row, a.pf set)
template <bool SYM>
static inline int up_cols_launch(const UpArgs &a, cudaStream_t st)
{
    constexpr int E = UpT<SYM>::E;
    const int npairs = a.n_in / 2;
    const int l2R = a.l2n - a.l2row;
    const unsigned grid_cols = (unsigned)(((size_t)a.B * npairs * E << a.l2row) / 256);
    switch (l2R) {
    case 1: return up_launch(k_up_cols<2, SYM>, a, grid_cols, 256, 0, st, "tree_up_cols");
    case 2: return up_launch(k_up_cols<4, SYM>, a, grid_cols, 256, 0, st, "tree_up_cols");
    case 3: return up_launch(k_up_cols<8, SYM>, a, grid_cols, 256, 0, st, "tree_up_cols");
    case 4: return up_launch(k_up_cols<16, SYM>, a, grid_cols, 256, 0, st, "tree_up_cols");
    case 5: return up_launch(k_up_cols<32, SYM>, a, grid_cols, 256, 0, st, "tree_up_cols");
    default:
        if (!a.last)
            return -1064;  // radix 64 only exists for the last level (coefficient output)
        return up_launch(k_up_cols<64, SYM>, a, grid_cols * 2, 128, 0, st, "tree_up_cols");
    }
}

// pending_cols (SYM, last level, row split only): the column pass is NOT launched; *pending_cols receives its
// arguments so that the caller can fuse it with the first chirp-z stage (chirpz2.cuh: k_up_cols_cz) or run it
// later with up_cols_pending
template <bool SYM>
static inline int up_level_t(UpArgs a, int l2smem_max, cudaStream_t st, UpArgs *pending_cols = nullptr)
{
    constexpr int E = UpT<SYM>::E;
    const int npairs = a.n_in / 2;
    const int N = 1 << a.l2n;
    cudaMemsetAsync(a.mx_out, 0, sizeof(double) * (size_t)a.B * npairs, st);
    static const char *names_s[3] = {"tree_up_smem_N2048", "tree_up_smem_N4096", "tree_up_smem_N8192"};
    if (a.l2n <= l2smem_max) {
        const unsigned grid = (unsigned)a.B * (unsigned)npairs * (unsigned)E;
        const size_t smem = sizeof(cplx) * N + 64 * sizeof(double);
        a.pf = up_pf_distance(UP_PF_SMEM11 + (a.l2n - 11));
        static const int knob_cluster = [] {
            const char *e = getenv("FNFT_B200_UP13_CLUSTER");  // 0: one CTA per item (k_up_smem<13>)
            return (e && e[0]) ? atoi(e) : 1;                  // measured 4.84 -> 4.54 ms per 4096 signals
        }();
        static const int knob_h13 = [] {  // positions per CTA of the N = 8192 cluster kernel: 12 (cluster of 2) or 11 (of 4)
            const char *e = getenv("FNFT_B200_UP13_L2H");
            return (e && e[0]) ? atoi(e) : 12;
        }();
        static const int knob_cluster12 = [] {  // 1: N = 4096 as a cluster of two CTAs of 2048 positions
            const char *e = getenv("FNFT_B200_UP12_CLUSTER");
            return (e && e[0]) ? atoi(e) : 0;
        }();
        if (a.l2n == 13 && !a.last && knob_cluster) {
            if (knob_h13 == 11)
                return up_launch_cluster(k_up_smem_cluster<13, 11, SYM>, a, grid, 4, st, names_s[2], 2048);
            return up_launch_cluster(k_up_smem_cluster<13, 12, SYM>, a, grid, 2, st, names_s[2]);
        }
        if (a.l2n == 12 && !a.last && knob_cluster12)
            return up_launch_cluster(k_up_smem_cluster<12, 11, SYM>, a, grid, 2, st, names_s[1], 2048);
        switch (a.l2n) {
        case 11: return up_launch(k_up_smem<11, SYM>, a, grid, 2048 / FNFTB_UP_TPP(11), smem, st, names_s[0]);
        case 12: return up_launch(k_up_smem<12, SYM>, a, grid, 4096 / FNFTB_UP_TPP(12), smem, st, names_s[1]);
        default: return up_launch(k_up_smem<13, SYM>, a, grid, 8192 / FNFTB_UP_TPP(13), smem, st, names_s[2]);
        }
    }
    static const int knob_cluster14 = [] {
        const char *e = getenv("FNFT_B200_UP14_CLUSTER");  // 0: row-split kernels (rows_a / cols / rows_c)
        return (e && e[0]) ? atoi(e) : 1;
    }();
    if (a.l2n == 14 && !a.last && knob_cluster14 && l2smem_max >= 13) {
        a.pf = up_pf_distance(UP_PF_SMEM13);
        return up_launch_cluster(k_up_smem_cluster<14, 12, SYM>, a, (unsigned)a.B * (unsigned)npairs * (unsigned)E, 4, st,
                                 "tree_up_smem_N16384");
    }
    a.l2row = FNFTB_UP_ROW_L2;
    const int l2R = a.l2n - a.l2row;
    const unsigned grid_rows = (unsigned)a.B * (unsigned)npairs * (unsigned)E << l2R;
    const size_t smem = sizeof(cplx) << a.l2row;
    if (SYM && a.last && pending_cols != nullptr && l2R <= 5) {
        // the whole level stays pending: the chirp-z either runs it as one cluster kernel (chirpz2.cuh:
        // k_up_last_cluster_cz) or as k_up_rows_a + k_up_cols_cz, and up_cols_pending runs the plain kernels
        a.rows_pending = 1;
        a.pf_rows = up_pf_distance(UP_PF_ROWS_A);
        a.pf = up_pf_distance(UP_PF_COLS);
        *pending_cols = a;
        return 0;
    }
    a.pf = up_pf_distance(UP_PF_ROWS_A);
    int rc = up_launch(k_up_rows_a<128, SYM>, a, grid_rows, 128, smem, st, "tree_up_rows_a");
    if (rc)
        return rc;
    a.pf = up_pf_distance(UP_PF_COLS);
    rc = up_cols_launch<SYM>(a, st);
    if (rc || a.last)
        return rc;
    a.pf = up_pf_distance(UP_PF_ROWS_C);
    return up_launch(k_up_rows_c<128, SYM>, a, grid_rows, 128, smem, st, "tree_up_rows_c");
}

int up_level(const UpArgs &a, int l2smem_max, cudaStream_t st, bool sym, UpArgs *pending_cols)
{
    return sym ? up_level_t<true>(a, l2smem_max, st, pending_cols) : up_level_t<false>(a, l2smem_max, st);
}
// k_up_rows_a of a pending last level (SYM)
int up_rows_a_pending(UpArgs &a, cudaStream_t st)
{
    if (!a.rows_pending)
        return 0;
    UpArgs r = a;
    r.pf = a.pf_rows;
    const unsigned grid_rows = (unsigned)a.B * (unsigned)(a.n_in / 2) * 2u << (a.l2n - a.l2row);
    a.rows_pending = 0;
    return up_launch(k_up_rows_a<128, true>, r, grid_rows, 128, sizeof(cplx) << a.l2row, st, "tree_up_rows_a");
}
int up_cols_pending(const UpArgs &a_in, cudaStream_t st)
{
    UpArgs a = a_in;
    const int rc = up_rows_a_pending(a, st);
    return rc ? rc : up_cols_launch<true>(a, st);
}
#else
int up_rows_a_pending(UpArgs &a, cudaStream_t st);
int up_level(const UpArgs &a, int l2smem_max, cudaStream_t st, bool sym = true, UpArgs *pending_cols = nullptr);
int up_cols_pending(const UpArgs &a, cudaStream_t st);
#endif
#endif  // !FNFTB_EMUL

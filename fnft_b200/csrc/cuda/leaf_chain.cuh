// fnft_b200 -- leaf construction for the higher-order exponential splittings.
//
// Replaces the remaining cases of fnft__akns_fscatter
//   /root/reference/src/private/fnft__akns_fscatter.c:256-400   (2SPLIT3A/3B/3S/4A, 4SPLIT4A)
//   /root/reference/src/private/fnft__akns_fscatter.c:435-912   (2SPLIT5A ... 2SPLIT8B)
// The reference spells every coefficient of every scheme out by hand.  Here the schemes are
// generated from their definition (Prins & Wahls, ICASSP 2018): with A = diag(-i*lambda,
// i*lambda) and B = [0 q; r 0], every scheme is a weighted sum of "chains"
//
//     C(m, first) = X(1/m) Y(2/m) X(2/m) ... (m+1 factors, sizes 1,2,2,...,2,1 over m)
//
// alternating X = exp(A.) and Y = exp(B.) when first = A (the "-A" schemes), or the other way
// round (the "-B" schemes).  m even is m/2 Strang steps, m odd ends on a half Strang step.
// A scheme of order p uses m in {1,3,..} (p odd) or {2,4,..} (p even) with the Richardson
// weights  c_k = prod_{j != k} m_k^2 / (m_k^2 - m_j^2), which cancel the error terms in
// 1/m^2.  2SPLIT3S (Burstein-Mirin) is 2/3 (C(2,A) + C(2,B)) - 1/6 (C(1,A) + C(1,B)).
//
// With z = exp(2 i lambda eps/deg), exp(A f eps) is diag(1, z^(f deg)) up to a scalar that the
// reference drops as well, so a chain is a sum over the 2^(nB+1) state sequences
// (row index before / after each exp(B.) factor): weight = product of the entries of the
// exp(B.) factors along the sequence, power of z = sum of the exp(A.) sizes met in state 2.
// At most 64 sequences of <= 5 factors per chain: one thread per sample enumerates them and
// accumulates into its own leaf in the level-0 buffer.
#pragma once
#include "tree_kernels.cuh"

#define FNFTB_CHAIN_MAX_TERMS 4

struct ChainTerm {
    double c;    // weight of the chain
    int m;       // number of half steps (the chain has m + 1 factors)
    int bfirst;  // 0: leftmost factor is exp(A.), 1: exp(B.)
};

struct ChainScheme {
    int nterms;
    int deg;  // degree of the leaf polynomials (every exp(A.) size times deg is an integer)
    ChainTerm t[FNFTB_CHAIN_MAX_TERMS];
};

// Scheme table keyed by fnft__akns_discretization_t
// (include/private/fnft__akns_discretization_t.h:43-72).  Returns 0 if `scheme` is not one
// of the chain schemes handled here.
static inline int chain_scheme_for(int scheme, ChainScheme *cs)
{
    int order = 0, bfirst = 0, deg = 0;
    switch (scheme) {
    case 6: order = 3; bfirst = 0; deg = 3; break;     // 2SPLIT3A
    case 7: order = 3; bfirst = 1; deg = 3; break;     // 2SPLIT3B
    case 9: order = 4; bfirst = 0; deg = 4; break;     // 2SPLIT4A
    case 20: order = 4; bfirst = 0; deg = 4; break;    // 4SPLIT4A
    case 11: order = 5; bfirst = 0; deg = 15; break;   // 2SPLIT5A
    case 12: order = 5; bfirst = 1; deg = 15; break;   // 2SPLIT5B
    case 13: order = 6; bfirst = 0; deg = 12; break;   // 2SPLIT6A
    case 14: order = 6; bfirst = 1; deg = 6; break;    // 2SPLIT6B
    case 15: order = 7; bfirst = 0; deg = 105; break;  // 2SPLIT7A
    case 16: order = 7; bfirst = 1; deg = 105; break;  // 2SPLIT7B
    case 17: order = 8; bfirst = 0; deg = 24; break;   // 2SPLIT8A
    case 18: order = 8; bfirst = 1; deg = 12; break;   // 2SPLIT8B
    default: return 0;
    }
    cs->deg = deg;
    cs->nterms = (order + 1) / 2;
    for (int k = 0; k < cs->nterms; ++k) {
        const int mk = (order & 1) ? 2 * k + 1 : 2 * k + 2;
        double c = 1.0;
        for (int j = 0; j < cs->nterms; ++j) {
            const int mj = (order & 1) ? 2 * j + 1 : 2 * j + 2;
            if (j != k)
                c *= (double)(mk * mk) / (double)(mk * mk - mj * mj);
        }
        cs->t[k].c = c;
        cs->t[k].m = mk;
        cs->t[k].bfirst = bfirst;
    }
    return 1;
}

struct LeafChainArgs {
    LeafArgs la;
    ChainScheme cs;
    int dpad;  // degree the leaves are stored with: next power of two >= cs.deg (see below)
};

// Degree padding.  The pair-product kernels are at their best when every level has a
// power-of-two degree d (cyclic length N = 2d with the analytic wrap correction, tree_kernels.cuh).
// A leaf of degree deg is therefore stored as z^(dpad-deg) * leaf, i.e. as a degree-dpad
// polynomial whose dpad-deg lowest coefficients are zero.  Like the z^deg*I padding matrices of
// fnft__poly_fmult.c:422-438 this only appends zeros BEHIND the deg*D+1 wanted coefficients
// (highest power first), which blk_tree_final strips.
static inline int chain_padded_degree(int deg)
{
    int d = 1;
    while (d < deg)
        d *= 2;
    return d;
}

// Adds c * C(m, first) for one sample to the four polynomials at o (d1 >= deg+1 slots each,
// highest power first: z^deg goes to slot 0).
HD void chain_accumulate(cplx *o, int d1, int deg, const ChainTerm &t, double eps_t, cplx q, cplx r)
{
    const int m = t.m;
    // the two distinct exp(B.) factors: sizes 1/m and 2/m  (fnft__akns_fscatter.c:46-59)
    cplx e[2][3];
    zero_freq_expm(e[0], eps_t / m, q, r);
    zero_freq_expm(e[1], 2.0 * eps_t / m, q, r);
    // factor i (0..m) is exp(B.) iff (i odd) == (leftmost is exp(A.))
    const int nB = t.bfirst ? (m + 2) / 2 : (m + 1) / 2;
    for (int mask = 0; mask < (2 << nB); ++mask) {
        int cur = mask & 1, ib = 0, expo = 0;
        cplx w = make_cplx(t.c, 0.0);
        for (int i = 0; i <= m; ++i) {
            const int size = (i == 0 || i == m) ? 1 : 2;
            const bool isB = ((i & 1) != 0) != (t.bfirst != 0);
            if (isB) {
                ++ib;
                const int nxt = (mask >> ib) & 1;
                const cplx f = (cur == nxt) ? e[size - 1][0] : (cur == 0 ? e[size - 1][1] : e[size - 1][2]);
                w = cmul(w, f);
                cur = nxt;
            } else if (cur) {
                expo += size * deg / m;  // power of z of this exp(A.) factor (an integer)
            }
        }
        cplx *dst = o + ((mask & 1) * 2 + cur) * d1 + (deg - expo);
        dst->x += w.x;
        dst->y += w.y;
    }
}

BLK void blk_leaf_chain(const LeafChainArgs &ca, blk3 bid, int nt, void * /*smem*/)
{
    const LeafArgs &a = ca.la;
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        const long long total = (long long)a.B * a.npad;
        if (gid < total) {
            const int s = (int)(gid / a.npad);
            const int m = (int)(gid % a.npad);
            const int d1 = ca.dpad + 1;
            cplx *o = a.out + (size_t)gid * 4 * d1;
            for (int i = 0; i < 4 * d1; ++i)
                o[i] = czero();
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
                for (int k = 0; k < ca.cs.nterms; ++k)
                    chain_accumulate(o, d1, a.deg0, ca.cs.t[k], a.eps_t, q, r);
            } else {
                // padding: z^deg * I (fnft__poly_fmult.c:422-438)
                o[0] = make_cplx(1.0, 0.0);
                o[3 * d1] = make_cplx(1.0, 0.0);
            }
            a.mx[gid] = 1.0;
        }
    }
}

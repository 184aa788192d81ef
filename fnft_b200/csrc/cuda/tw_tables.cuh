// fnft_b200 -- pass-major twiddle tables for the spectrum-carry tree kernels
// (tree_low2.cuh, tree_up.cuh).
//
// A radix-R pass at stride s multiplies element q of butterfly o by w_len^(q*o), len = R*s.
// Reading those factors from one master table exp(-2*pi*i*k/4096) makes the lanes of a warp
// (consecutive o) hit addresses q*4096/len entries apart -- one 32-byte sector per lane.
// Here every (len, R) has its own table laid out [q-1][o], so a warp reads 512 contiguous
// bytes per factor; likewise the "twist" factors w_2N^i, i < N, of every operand length N.
// All entries are exact-argument sincospi values (the reference recomputes sin/cos per plan,
// src/3rd_party/kiss_fft/kiss_fft.c:357-363).
#pragma once
#ifndef FNFTB_EMUL
#include "common.cuh"

#define FNFTB_TW_MINL 3
#define FNFTB_TW_MAXL 18  // tables for lengths 2^3 .. 2^18

struct TwSet {
    const cplx *base;
    int twist_off[FNFTB_TW_MAXL + 1];    // [log2 N]      -> w_2N^i, i < N
    int pass_off[FNFTB_TW_MAXL + 1][7];  // [log2 len][log2 R] -> [q-1][o], o < len/R (radix 64: longest length only)
};

static __global__ void k_tw_fill_pass(cplx *dst, int l2len, int l2r)
{
    const int len = 1 << l2len, s = len >> l2r;
    const int total = ((1 << l2r) - 1) * s;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < total) {
        const int q = idx / s + 1, o = idx % s;
        double sn, cs;
        sincospi(-2.0 * (double)(q * o) / (double)len, &sn, &cs);
        dst[idx] = make_cplx(cs, sn);
    }
}

static __global__ void k_tw_fill_twist(cplx *dst, int l2n)
{
    const int N = 1 << l2n;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N) {
        double sn, cs;
        sincospi(-(double)idx / (double)N, &sn, &cs);
        dst[idx] = make_cplx(cs, sn);
    }
}

// number of table entries; fills the offsets of *T (base left untouched)
static inline size_t twset_layout(TwSet *T)
{
    size_t off = 0;
    for (int l = 0; l <= FNFTB_TW_MAXL; ++l) {
        T->twist_off[l] = -1;
        for (int r = 0; r < 7; ++r)
            T->pass_off[l][r] = -1;
    }
    for (int l = FNFTB_TW_MINL; l <= FNFTB_TW_MAXL; ++l) {
        T->twist_off[l] = (int)off;
        off += (size_t)1 << l;
        for (int r = 1; r <= (l == FNFTB_TW_MAXL ? 6 : 5) && r <= l - 2; ++r) {
            T->pass_off[l][r] = (int)off;
            off += (size_t)((1 << r) - 1) * ((size_t)1 << (l - r));
        }
    }
    return off;
}

static inline int twset_build(TwSet *T, cplx *mem, cudaStream_t st)
{
    twset_layout(T);
    T->base = mem;
    for (int l = FNFTB_TW_MINL; l <= FNFTB_TW_MAXL; ++l) {
        const int N = 1 << l;
        k_tw_fill_twist<<<(N + 255) / 256, 256, 0, st>>>(mem + T->twist_off[l], l);
        for (int r = 1; r <= (l == FNFTB_TW_MAXL ? 6 : 5) && r <= l - 2; ++r) {
            const int total = ((1 << r) - 1) * (N >> r);
            k_tw_fill_pass<<<(total + 255) / 256, 256, 0, st>>>(mem + T->pass_off[l][r], l, r);
        }
    }
    return (int)cudaGetLastError();
}
#endif

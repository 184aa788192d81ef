// fnft_b200 -- coefficient form -> spectrum-carry form of a tree level (general 2x2 case).
//
// The degree >= 3 splitting schemes (leaf_chain.cuh) start the product tree in coefficient form
// with the pair-product kernels of tree_kernels.cuh, whose row-split stage ends at cyclic length
// 65536.  Once the matrices have degree 1024 this kernel turns every entry into its values at the
// 2048-th roots of unity (bit-reversed order) plus the top / bottom coefficients, which is the
// input format of the spectrum-carry upper levels (tree_up.cuh, E = 4): those cost a third less
// arithmetic per level and never form coefficients in HBM again, and reach final degree 131072.  The lazy normalisation carries over unchanged (values and tops are those of the
// unscaled coefficients; mx[] keeps max|c|).
//
//   V[k] = sum_{i=0}^{d} c_i w_N^(ik),  N = 2d = 2048:  FFT_N of c_0..c_{d-1} (zero padded)
//          plus c_d (-1)^k; position p of the output holds bin bitrev(p), so (-1)^k = -1 for p >= N/2.
#pragma once
#ifndef FNFTB_EMUL
#include "tree_up.cuh"

struct ConvArgs {
    const cplx *in;   // [B*n][4][d+1] coefficients, d = 1024
    cplx *out;        // [B*n][4][N] values, N = 2048
    GenTops *tt_out;  // [B*n]
    TwSet tw;
    long long nmat;   // B*n
};

// grid.x = nmat * 4, 128 threads, 32 KiB shared memory
__global__ void __launch_bounds__(128, 4) k_coef_to_spec2048(const ConvArgs a)
{
    constexpr int N = 2048, d = 1024, NT = 128;
    extern __shared__ double2 fnftb_smem_conv[];
    cplx *S = (cplx *)fnftb_smem_conv;
    const int tid = threadIdx.x;
    const int e = blockIdx.x & 3;
    const size_t mat = blockIdx.x >> 2;
    const cplx *c = a.in + (mat * 4 + e) * (size_t)(d + 1);
    for (int i = tid; i < N; i += NT)
        S[swz2(i)] = (i < d) ? c[i] : czero();
    const cplx ctop = c[d];
    if (tid == 0) {
        a.tt_out[mat].t[e] = ctop;
        a.tt_out[mat].b[e] = c[0];
    }
    __syncthreads();
    up_p_pass<16, -1>(S, N, 7, a.tw, tid, NT);
    __syncthreads();
    up_p_pass<16, -1>(S, N, 3, a.tw, tid, NT);
    __syncthreads();
    // last pass: radix 8 at stride 1 (no twiddles), straight to global memory with the top term
    cplx *o = a.out + (mat * 4 + e) * (size_t)N;
    for (int g = tid; g < N / 8; g += NT) {
        const int base = g << 3;
        cplx v[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2)
            v[n2] = S[swz2(base + n2)];
        Dft<8, -1>::run(v);
        const double sg = (base >= N / 2) ? -1.0 : 1.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pos = base + brev_c(q, 3);
            o[pos] = make_cplx(v[q].x + sg * ctop.x, v[q].y + sg * ctop.y);
        }
    }
}

static inline int coef_to_spec2048(const ConvArgs &a, cudaStream_t st)
{
    auto kern = k_coef_to_spec2048;
    if (g_fnftb_profile_on)
        fnftb_profile_begin("tree_coef_to_spec", st);
    kern<<<(unsigned)(a.nmat * 4), 128, sizeof(cplx) * 2048, st>>>(a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(st);
    ++g_fnftb_launch_count;
    return (int)cudaGetLastError();
}
#endif

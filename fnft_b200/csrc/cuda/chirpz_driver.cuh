// fnft_b200 -- host-side driver of the batched chirp-z evaluation.
#pragma once
#include "chirpz_kernels.cuh"
#include "launch.cuh"

#ifndef FNFTB_CZ_ROW_N
#define FNFTB_CZ_ROW_N 4096
#endif

struct CzGeom {
    int L, N1, N2, C;
};

static inline CzGeom cz_geometry(int deg, int M, int row_n = FNFTB_CZ_ROW_N)
{
    CzGeom g;
    size_t need = (size_t)deg + (size_t)M;  // N + M - 1 with N = deg + 1
    size_t L = 1;
    while (L < need)
        L *= 2;
    g.L = (int)L;
    g.N2 = (int)(L < (size_t)row_n ? L : (size_t)row_n);
    g.N1 = g.L / g.N2;
    int C = 2048 / g.N1;
    if (C < 1)
        C = 1;
    if (C > g.N2)
        C = g.N2;
    if (C > 128)
        C = 128;
    g.C = C;
    return g;
}

// elements of the signal-independent tables (tab_y | tab_out | tab_ph | tab_tw)
static inline size_t cz_table_elems(const CzGeom &g, int deg, int M)
{
    return (size_t)(deg + 1) + (size_t)M + 3 * (size_t)M + (size_t)g.L + 8;
}

// elements needed in ybuf for B signals / in vhat
static inline size_t cz_ybuf_elems(const CzGeom &g, size_t B, int npoly) { return B * npoly * (size_t)g.L; }

// (instantiated in k_chirpz2.cu and in the host emulation only)
#if defined(FNFTB_TU_CZ2) || defined(FNFTB_EMUL)
// Fills geometry/plans in `a` (a.deg, a.M, a.B, a.npoly must be set) and runs the
// whole evaluation: vhat, forward columns, rows, inverse columns + epilogue.
// `tables` must provide cz_table_elems() elements of device scratch.
static inline int cz_run(CzArgs a, cplx *tables, fnftb_stream_t st, int row_n = FNFTB_CZ_ROW_N)
{
    const CzGeom g = cz_geometry(a.deg, a.M, row_n);
    a.L = g.L;
    a.N1 = g.N1;
    a.N2 = g.N2;
    a.C = g.C;
    a.log2C = ilog2i((unsigned)g.C);
    a.plan1 = make_fft_plan(g.N1);
    a.plan2 = make_fft_plan(g.N2);
    const int nt = 256;
    int rc;
    a.tab_y = tables;
    a.tab_out = a.tab_y + (a.deg + 1);
    a.tab_ph = a.tab_out + a.M;
    a.tab_tw = a.tab_ph + 3 * (size_t)a.M;
    // 0. signal-independent tables (chirps, four-step twiddles, epilogue phases)
    {
        long long tot = a.deg + 1;
        if (a.M > tot)
            tot = a.M;
        if (g.L > tot)
            tot = g.L;
        rc = launch_blocks<CzArgs, blk_cz_tables>(a, (unsigned)((tot + nt - 1) / nt), nt, 0, st,
                                                  "cz_filter");
        if (rc)
            return rc;
    }
    // 1. spectrum of the chirp filter
    {
        CzArgs v = a;
        v.gen_v = 1;
        v.fwd_only = 1;
        rc = launch_blocks<CzArgs, blk_cz_cols_fwd, 256, 3>(v, (unsigned)(g.N2 / g.C), nt,
                                                    cz_cols_smem_bytes(g.C, g.N1, 1), st, "cz_filter");
        if (rc)
            return rc;
        rc = launch_blocks<CzArgs, blk_cz_rows, 256, 2>(v, (unsigned)g.N1, nt, sizeof(cplx) * (size_t)g.N2, st,
                                                "cz_filter");
        if (rc)
            return rc;
    }
    a.gen_v = 0;
    a.fwd_only = 0;
    // 2. forward columns of all polynomials
    rc = launch_blocks<CzArgs, blk_cz_cols_fwd, 256, 3>(a, (unsigned)((size_t)a.B * a.npoly * (g.N2 / g.C)), nt,
                                                cz_cols_smem_bytes(g.C, g.N1, 1), st, "cz_cols_fwd");
    if (rc)
        return rc;
    // 3. rows: FFT, multiply, inverse FFT
    rc = launch_blocks<CzArgs, blk_cz_rows, 256, 2>(a, (unsigned)((size_t)a.B * a.npoly * g.N1), nt,
                                            sizeof(cplx) * (size_t)g.N2, st, "cz_rows");
    if (rc)
        return rc;
    // 4. inverse columns + epilogue
    return launch_blocks<CzArgs, blk_cz_cols_inv, 256, 3>(a, (unsigned)((size_t)a.B * (g.N2 / g.C)), nt,
                                                  cz_cols_smem_bytes(g.C, g.N1, a.npoly), st, "cz_cols_inv");
}
#endif
